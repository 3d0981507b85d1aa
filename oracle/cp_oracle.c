/* TEST INFRASTRUCTURE ONLY -- parity oracle, never the product path (see cp_oracle.h).
 *
 * Plain-C restatement of the CPU critical-point tracker of hguo/ftk (non-GMP build).
 * Citations "ref: <path>:<lines>" are relative to the reference checkout (/root/reference).
 * Compile with -fwrapv -ffp-contract=off (oracle/Makefile): the reference relies on
 * two's-complement wrap of int64 determinants and x86-64 has no FMA contraction by default.
 */
#include "cp_oracle.h"

#include <limits.h>
#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * 1. implicit simplicial mesh tables     ref: include/ftk/mesh/simplicial_regular_mesh.hh
 * ---------------------------------------------------------------------------------------- */
#define MAXT 64
#define MAXV 5
#define MAXSO 96
#define MAXND 4

typedef struct { int nv; int v[MAXV][MAXND]; } simplex_t;
typedef struct { int type; int off[MAXND]; } tyoff_t;

typedef struct {
  int nd, ready;
  int ntypes[MAXND + 1];
  simplex_t unit[MAXND + 1][MAXT];
  int n_ord[MAXND + 1], n_int[MAXND + 1];
  int ord_types[MAXND + 1][MAXT], int_types[MAXND + 1][MAXT], is_ord[MAXND + 1][MAXT];
  int nsides[MAXND + 1][MAXT];  tyoff_t sides[MAXND + 1][MAXT][MAXV + 1];
  int nsideof[MAXND + 1][MAXT]; tyoff_t sideof[MAXND + 1][MAXT][MAXSO];
} mesh_t;

static int cmp_vertex(const int *a, const int *b, int nd)
{ /* std::vector<int> lexicographic compare */
  for (int i = 0; i < nd; i ++) if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return 0;
}

static int cmp_simplex(const simplex_t *a, const simplex_t *b, int nd)
{ /* std::vector<std::vector<int>> lexicographic compare */
  const int n = a->nv < b->nv ? a->nv : b->nv;
  for (int i = 0; i < n; i ++) {
    const int c = cmp_vertex(a->v[i], b->v[i], nd);
    if (c) return c;
  }
  return (a->nv > b->nv) - (a->nv < b->nv);
}

static void sort_vertices(simplex_t *s, int nd)
{ /* std::sort(simplex.begin(), simplex.end()) */
  for (int i = 1; i < s->nv; i ++)
    for (int j = i; j > 0 && cmp_vertex(s->v[j-1], s->v[j], nd) > 0; j --) {
      int tmp[MAXND];
      memcpy(tmp, s->v[j], sizeof(tmp)); memcpy(s->v[j], s->v[j-1], sizeof(tmp)); memcpy(s->v[j-1], tmp, sizeof(tmp));
    }
}

/* ref: simplicial_regular_mesh.hh:620-653 subdivide_unit_cube */
static int subdivide_unit_cube(int n, simplex_t *out /* n! entries */)
{
  if (n == 1) {
    out[0].nv = 2; memset(out[0].v, 0, sizeof(out[0].v));
    out[0].v[0][0] = 0; out[0].v[1][0] = 1;
    return 1;
  }
  simplex_t *r0 = (simplex_t *)malloc(sizeof(simplex_t) * 24);
  const int n0 = subdivide_unit_cube(n - 1, r0);
  int cnt = 0;
  for (int i = 0; i < n; i ++)
    for (int j = 0; j < n0; j ++) {
      simplex_t s; memset(&s, 0, sizeof(s));
      s.nv = n + 1;
      for (int k = 0; k < n; k ++) { /* vertex.insert(vertex.begin()+i, 0) */
        int p = 0;
        for (int q = 0; q < n; q ++) s.v[k][q] = (q == i) ? 0 : r0[j].v[k][p ++];
      }
      for (int q = 0; q < n; q ++) s.v[n][q] = 1; /* all_one_vertex */
      out[cnt ++] = s;
    }
  free(r0);
  return cnt;
}

/* ref: simplicial_regular_mesh.hh:655-683 reduce_unit_simplex */
static void reduce_unit_simplex(simplex_t *s, int nd, int *offset)
{
  for (int i = 0; i < nd; i ++) {
    int all_one = 1;
    for (int j = 0; j < s->nv; j ++) if (s->v[j][i] == 0) { all_one = 0; break; }
    offset[i] = 0;
    if (all_one) { offset[i] = 1; for (int j = 0; j < s->nv; j ++) s->v[j][i] = 0; }
  }
}

static int set_insert(simplex_t *set, int n, const simplex_t *s, int nd)
{ /* std::set<...>::insert keeping lexicographic order */
  int pos = 0;
  while (pos < n) {
    const int c = cmp_simplex(&set[pos], s, nd);
    if (c == 0) return n;
    if (c > 0) break;
    pos ++;
  }
  for (int i = n; i > pos; i --) set[i] = set[i-1];
  set[pos] = *s;
  return n + 1;
}

/* ref: simplicial_regular_mesh.hh:685-715 enumerate_unit_simplices.  The reference drops the
 * last vertex of every permutation of a (k+1)-simplex; that is the same set as dropping each
 * vertex in turn. */
static void enumerate_unit_simplices(mesh_t *m)
{
  const int n = m->nd;
  simplex_t cube[24];
  const int nc = subdivide_unit_cube(n, cube);
  m->ntypes[n] = 0;
  for (int i = 0; i < nc; i ++) {
    sort_vertices(&cube[i], n);
    m->ntypes[n] = set_insert(m->unit[n], m->ntypes[n], &cube[i], n);
  }
  for (int k = n - 1; k >= 0; k --) {
    m->ntypes[k] = 0;
    for (int h = 0; h < m->ntypes[k+1]; h ++) {
      const simplex_t *hs = &m->unit[k+1][h];
      for (int drop = 0; drop < hs->nv; drop ++) {
        simplex_t s; memset(&s, 0, sizeof(s));
        for (int j = 0; j < hs->nv; j ++) if (j != drop) { memcpy(s.v[s.nv], hs->v[j], sizeof(int) * MAXND); s.nv ++; }
        int off[MAXND];
        reduce_unit_simplex(&s, n, off);
        sort_vertices(&s, n);
        if (m->ntypes[k] >= MAXT - 1) { fprintf(stderr, "cp_oracle: MAXT exceeded\n"); abort(); }
        m->ntypes[k] = set_insert(m->unit[k], m->ntypes[k], &s, n);
      }
    }
  }
}

static int cmp_tyoff(const tyoff_t *a, const tyoff_t *b, int nd)
{ /* std::tuple<int, std::vector<int>> compare */
  if (a->type != b->type) return a->type < b->type ? -1 : 1;
  return cmp_vertex(a->off, b->off, nd);
}

static int tyoff_insert(tyoff_t *set, int n, const tyoff_t *e, int nd, int cap)
{
  int pos = 0;
  while (pos < n) {
    const int c = cmp_tyoff(&set[pos], e, nd);
    if (c == 0) return n;
    if (c > 0) break;
    pos ++;
  }
  if (n >= cap) { fprintf(stderr, "cp_oracle: side table capacity exceeded\n"); abort(); }
  for (int i = n; i > pos; i --) set[i] = set[i-1];
  set[pos] = *e;
  return n + 1;
}

/* ref: simplicial_regular_mesh.hh:760-797 enumerate_unit_simplex_sides */
static void enumerate_sides(mesh_t *m, int k, int type)
{
  m->nsides[k][type] = 0;
  if (k == 0) return;
  const simplex_t *s = &m->unit[k][type];
  for (int drop = 0; drop < s->nv; drop ++) {
    simplex_t side; memset(&side, 0, sizeof(side));
    for (int j = 0; j < s->nv; j ++) if (j != drop) { memcpy(side.v[side.nv], s->v[j], sizeof(int) * MAXND); side.nv ++; }
    tyoff_t e; memset(&e, 0, sizeof(e));
    reduce_unit_simplex(&side, m->nd, e.off);
    sort_vertices(&side, m->nd);
    for (int t = 0; t < m->ntypes[k-1]; t ++)
      if (cmp_simplex(&side, &m->unit[k-1][t], m->nd) == 0) {
        e.type = t;
        m->nsides[k][type] = tyoff_insert(m->sides[k][type], m->nsides[k][type], &e, m->nd, MAXV + 1);
      }
  }
}

/* ref: simplicial_regular_mesh.hh:717-758 enumerate_unit_simplex_side_of */
static void enumerate_side_of(mesh_t *m, int k, int type)
{
  m->nsideof[k][type] = 0;
  if (k == m->nd) return;
  const simplex_t *s = &m->unit[k][type]; /* at zero corner */
  int corner[MAXND];
  int ncorners = 1; for (int i = 0; i < m->nd; i ++) ncorners *= 3;
  for (int ci = 0; ci < ncorners; ci ++) {
    int c = ci;
    for (int i = 0; i < m->nd; i ++) { corner[i] = (c % 3) - 1; c /= 3; }
    for (int t = 0; t < m->ntypes[k+1]; t ++) {
      const simplex_t *h = &m->unit[k+1][t];
      int includes = 1;
      for (int a = 0; a < s->nv && includes; a ++) {
        int found = 0;
        for (int b = 0; b < h->nv && !found; b ++) {
          int eq = 1;
          for (int d = 0; d < m->nd; d ++) if (h->v[b][d] + corner[d] != s->v[a][d]) { eq = 0; break; }
          found = eq;
        }
        includes = found;
      }
      if (includes) {
        tyoff_t e; memset(&e, 0, sizeof(e));
        e.type = t; memcpy(e.off, corner, sizeof(int) * m->nd);
        m->nsideof[k][type] = tyoff_insert(m->sideof[k][type], m->nsideof[k][type], &e, m->nd, MAXSO);
      }
    }
  }
}

/* ref: simplicial_regular_mesh.hh:799-831 derive_ordinal_and_interval_simplices */
static void derive_ordinal_interval(mesh_t *m)
{
  for (int d = 0; d <= m->nd; d ++) {
    m->n_ord[d] = m->n_int[d] = 0;
    if (d == 0) { m->ord_types[0][m->n_ord[0] ++] = 0; m->is_ord[0][0] = 1; continue; }
    for (int t = 0; t < m->ntypes[d]; t ++) {
      int time = 0;
      for (int i = 0; i < m->unit[d][t].nv; i ++) time += m->unit[d][t].v[i][m->nd - 1];
      if (time == 0) { m->ord_types[d][m->n_ord[d] ++] = t; m->is_ord[d][t] = 1; }
      else { m->int_types[d][m->n_int[d] ++] = t; m->is_ord[d][t] = 0; }
    }
  }
}

static mesh_t g_mesh[2];

/* ref: simplicial_regular_mesh.hh:891-927 initialize_subdivision */
static const mesh_t *get_mesh(int nd_mesh)
{
  mesh_t *m = &g_mesh[nd_mesh - 3];
#ifdef _OPENMP
#pragma omp critical(cpo_mesh_init)
#endif
  {
    if (!m->ready) {
      memset(m, 0, sizeof(*m));
      m->nd = nd_mesh;
      enumerate_unit_simplices(m);
      for (int k = 0; k <= m->nd; k ++) for (int t = 0; t < m->ntypes[k]; t ++) enumerate_sides(m, k, t);
      for (int k = 0; k <= m->nd; k ++) for (int t = 0; t < m->ntypes[k]; t ++) enumerate_side_of(m, k, t);
      derive_ordinal_interval(m);
      m->ready = 1;
    }
  }
  return m;
}

int cpo_mesh_ntypes(int nd_mesh, int k, int scope)
{
  const mesh_t *m = get_mesh(nd_mesh);
  return scope == 0 ? m->ntypes[k] : scope == 1 ? m->n_ord[k] : m->n_int[k];
}

void cpo_mesh_unit_simplex(int nd_mesh, int k, int type, int32_t *out)
{
  const mesh_t *m = get_mesh(nd_mesh);
  for (int i = 0; i <= k; i ++) for (int j = 0; j < nd_mesh; j ++) out[i * nd_mesh + j] = m->unit[k][type].v[i][j];
}

int cpo_mesh_scope_type(int nd_mesh, int k, int scope, int itype)
{
  const mesh_t *m = get_mesh(nd_mesh);
  return scope == 0 ? itype : scope == 1 ? m->ord_types[k][itype] : m->int_types[k][itype];
}

int cpo_mesh_sides(int nd_mesh, int k, int type, int32_t *out)
{
  const mesh_t *m = get_mesh(nd_mesh);
  for (int i = 0; i < m->nsides[k][type]; i ++) {
    out[i * (nd_mesh + 1)] = m->sides[k][type][i].type;
    for (int j = 0; j < nd_mesh; j ++) out[i * (nd_mesh + 1) + 1 + j] = m->sides[k][type][i].off[j];
  }
  return m->nsides[k][type];
}

int cpo_mesh_side_of(int nd_mesh, int k, int type, int32_t *out)
{
  const mesh_t *m = get_mesh(nd_mesh);
  for (int i = 0; i < m->nsideof[k][type]; i ++) {
    out[i * (nd_mesh + 1)] = m->sideof[k][type][i].type;
    for (int j = 0; j < nd_mesh; j ++) out[i * (nd_mesh + 1) + 1 + j] = m->sideof[k][type][i].off[j];
  }
  return m->nsideof[k][type];
}

/* ------------------------------------------------------------------------------------------
 * 2. SoS predicates       ref: include/ftk/numeric/{sign,det,sign_det,critical_point_test}.hh
 * ---------------------------------------------------------------------------------------- */
typedef int64_t i64;

static int sgn(i64 x) { return (0 < x) - (x < 0); }                          /* ref: sign.hh:10-14 */
static i64 det2v(i64 a, i64 b) { return a * 1 - b * 1; }                     /* det2 of {{a,1},{b,1}}, ref: det.hh:11-16 */
static i64 det3(const i64 m[3][3])                                          /* ref: det.hh:18-26 */
{
  return m[0][0] * (m[1][1]*m[2][2] - m[1][2]*m[2][1])
       - m[0][1] * (m[1][0]*m[2][2] - m[1][2]*m[2][0])
       + m[0][2] * (m[1][0]*m[2][1] - m[1][1]*m[2][0]);
}
static i64 det3c(i64 a0, i64 a1, i64 b0, i64 b1, i64 c0, i64 c1)
{ /* det3 of {{a0,a1,1},{b0,b1,1},{c0,c1,1}} */
  const i64 m[3][3] = {{a0, a1, 1}, {b0, b1, 1}, {c0, c1, 1}};
  return det3(m);
}
static i64 det4(const i64 m[4][4])                                          /* ref: det.hh:28-55 */
{
  const i64
    d2233 = m[2][2] * m[3][3] - m[2][3] * m[3][2],
    d2133 = m[2][1] * m[3][3] - m[2][3] * m[3][1],
    d2132 = m[2][1] * m[3][2] - m[2][2] * m[3][1],
    d2033 = m[2][0] * m[3][3] - m[2][3] * m[3][0],
    d2032 = m[2][0] * m[3][2] - m[2][2] * m[3][0],
    d2031 = m[2][0] * m[3][1] - m[2][1] * m[3][0];
  return m[0][0] * (m[1][1] * d2233 - m[1][2] * d2133 + m[1][3] * d2132)
       - m[0][1] * (m[1][0] * d2233 - m[1][2] * d2033 + m[1][3] * d2032)
       + m[0][2] * (m[1][0] * d2133 - m[1][1] * d2033 + m[1][3] * d2031)
       - m[0][3] * (m[1][0] * d2132 - m[1][1] * d2032 + m[1][2] * d2031);
}

/* ref: sign_det.hh:44-90 robust_sign_det3 */
static int robust_sign_det3(const i64 X[3][2])
{
  int s;
  s = sgn(det3c(X[0][0], X[0][1], X[1][0], X[1][1], X[2][0], X[2][1])); if (s) return s;
  s = -sgn(det2v(X[1][0], X[2][0])); if (s) return s;
  s = sgn(det2v(X[1][1], X[2][1])); if (s) return s;
  s = sgn(det2v(X[0][0], X[2][0])); if (s) return s;
  return 1;
}

/* ref: sign_det.hh:92-200 robust_sign_det4 */
static int robust_sign_det4(const i64 X[4][3])
{
  int s;
  { const i64 M[4][4] = {{X[0][0], X[0][1], X[0][2], 1}, {X[1][0], X[1][1], X[1][2], 1},
                         {X[2][0], X[2][1], X[2][2], 1}, {X[3][0], X[3][1], X[3][2], 1}};
    s = sgn(det4(M)); if (s) return s; }                                                          /* t=0 */
  s =  sgn(det3c(X[1][0], X[1][1], X[2][0], X[2][1], X[3][0], X[3][1])); if (s) return s;         /* t=1 */
  s = -sgn(det3c(X[1][0], X[1][2], X[2][0], X[2][2], X[3][0], X[3][2])); if (s) return s;         /* t=2 */
  s =  sgn(det3c(X[1][1], X[1][2], X[2][1], X[2][2], X[3][1], X[3][2])); if (s) return s;         /* t=3 */
  s = -sgn(det3c(X[0][0], X[0][1], X[2][0], X[2][1], X[3][0], X[3][1])); if (s) return s;         /* t=4 */
  s =  sgn(det2v(X[2][0], X[3][0])); if (s) return s;                                             /* t=5 */
  s = -sgn(det2v(X[2][1], X[3][1])); if (s) return s;                                             /* t=6 */
  s =  sgn(det3c(X[0][0], X[0][2], X[2][0], X[2][2], X[3][0], X[3][2])); if (s) return s;         /* t=7 */
  s =  sgn(det2v(X[2][2], X[3][2])); if (s) return s;                                             /* t=8 */
  s = -sgn(det3c(X[0][1], X[0][2], X[2][1], X[2][2], X[3][1], X[3][2])); if (s) return s;         /* t=9 */
  s =  sgn(det3c(X[0][0], X[0][1], X[1][0], X[1][1], X[3][0], X[3][1])); if (s) return s;         /* t=10 */
  s = -sgn(det2v(X[1][0], X[3][0])); if (s) return s;                                             /* t=11 */
  s =  sgn(det2v(X[1][1], X[3][1])); if (s) return s;                                             /* t=12 */
  s =  sgn(det2v(X[0][0], X[3][0])); if (s) return s;                                             /* t=13 */
  return 1;
}

/* ref: sign_det.hh:203-220 nswaps_bubble_sort */
static int nswaps_bubble_sort(int n, int *arr, int *order)
{
  for (int i = 0; i < n; i ++) order[i] = i;
  int nswaps = 0;
  for (int i = 0; i < n - 1; i ++)
    for (int j = 0; j < n - i - 1; j ++)
      if (arr[j] > arr[j+1]) {
        int t = arr[j]; arr[j] = arr[j+1]; arr[j+1] = t;
        t = order[j]; order[j] = order[j+1]; order[j+1] = t;
        nswaps ++;
      }
  return nswaps;
}

/* ref: sign_det.hh:243-266 positive2 */
int cpo_positive2(const int64_t X1[3][2], const int32_t indices1[3])
{
  int indices[3], orders[3];
  for (int i = 0; i < 3; i ++) indices[i] = indices1[i];
  const int s = nswaps_bubble_sort(3, indices, orders);
  i64 X[3][2];
  for (int i = 0; i < 3; i ++) for (int j = 0; j < 2; j ++) X[i][j] = X1[orders[i]][j];
  int d = robust_sign_det3(X);
  if (s % 2 != 0) d = -d;
  return d;
}

/* ref: sign_det.hh:268-289 positive3 */
int cpo_positive3(const int64_t X1[4][3], const int32_t indices1[4])
{
  int indices[4], orders[4];
  for (int i = 0; i < 4; i ++) indices[i] = indices1[i];
  const int s = nswaps_bubble_sort(4, indices, orders);
  i64 X[4][3];
  for (int i = 0; i < 4; i ++) for (int j = 0; j < 3; j ++) X[i][j] = X1[orders[i]][j];
  int d = robust_sign_det4(X);
  if (s % 2 != 0) d = -d;
  return d;
}

/* ref: sign_det.hh:360-388 robust_point_in_simplex2 with x = 0, ix = -1
 *      (critical_point_test.hh:22-27 robust_critical_point_in_simplex2) */
int cpo_robust_cp_in_simplex2(const int64_t X[3][2], const int32_t indices[3])
{
  const int s = cpo_positive2(X, indices);
  for (int i = 0; i < 3; i ++) {
    i64 Y[3][2]; int32_t my[3];
    for (int j = 0; j < 3; j ++)
      if (i == j) { my[j] = -1; Y[j][0] = Y[j][1] = 0; }
      else { my[j] = indices[j]; Y[j][0] = X[j][0]; Y[j][1] = X[j][1]; }
    if (s != cpo_positive2(Y, my)) return 0;
  }
  return 1;
}

/* ref: sign_det.hh:390-414 robust_point_in_simplex3; critical_point_test.hh:29-34 */
int cpo_robust_cp_in_simplex3(const int64_t X[4][3], const int32_t indices[4])
{
  const int s = cpo_positive3(X, indices);
  for (int i = 0; i < 4; i ++) {
    i64 Y[4][3]; int32_t my[4];
    for (int j = 0; j < 4; j ++)
      if (i == j) { my[j] = -1; Y[j][0] = Y[j][1] = Y[j][2] = 0; }
      else { my[j] = indices[j]; for (int k = 0; k < 3; k ++) Y[j][k] = X[j][k]; }
    if (s != cpo_positive3(Y, my)) return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------
 * 3. floating-point leaf functions
 * ---------------------------------------------------------------------------------------- */
static double dmax(double a, double b) { return (a < b) ? b : a; }  /* std::max(a, b) */
static double dmin(double a, double b) { return (b < a) ? b : a; }  /* std::min(a, b) */

/* ref: include/ftk/numeric/clamp.hh:15-37 */
static void clamp_barycentric(int n, double *x)
{
  double sum = 0.0;
  for (int i = 0; i < n; i ++) { x[i] = dmin(dmax(0.0, x[i]), 1.0); sum += x[i]; }
  for (int i = 0; i < n; i ++) x[i] /= sum;
  if (isnan(x[0]) || isinf(x[0])) for (int i = 0; i < n; i ++) x[i] = 1.0 / n;
}

/* ref: include/ftk/numeric/inverse_linear_interpolation_solver.hh:32-54 (cond is dead code)
 *      include/ftk/numeric/linear_solver.hh:22-34 solve_linear2x2 (Cramer) */
static int inverse_lerp_s2v2(const double V[3][2], double mu[3])
{
  const double eps = DBL_EPSILON;
  const double A[2][2] = {{V[0][0] - V[2][0], V[1][0] - V[2][0]}, {V[0][1] - V[2][1], V[1][1] - V[2][1]}};
  const double b[2] = {-V[2][0], -V[2][1]};
  const double D  = A[0][0] * A[1][1] - A[1][0] * A[0][1],
               Dx = b[0] * A[1][1] - A[0][1] * b[1],
               Dy = A[0][0] * b[1] - b[0] * A[1][0];
  mu[0] = Dx / D; mu[1] = Dy / D;
  mu[2] = 1.0 - mu[0] - mu[1];
  return mu[0] >= -eps && mu[0] <= 1.0 + eps && mu[1] >= -eps && mu[1] <= 1.0 + eps && mu[2] >= -eps && mu[2] <= 1.0 + eps;
}

/* ref: inverse_linear_interpolation_solver.hh:143-167; linear_solver.hh:12-20;
 *      matrix_inverse.hh:23-45; matrix_multiplication.hh:53-60 */
static int inverse_lerp_s3v3(const double V[4][3], double l[4])
{
  const double eps = DBL_EPSILON;
  const double m[3][3] = {
    {V[0][0] - V[3][0], V[1][0] - V[3][0], V[2][0] - V[3][0]},
    {V[0][1] - V[3][1], V[1][1] - V[3][1], V[2][1] - V[3][1]},
    {V[0][2] - V[3][2], V[1][2] - V[3][2], V[2][2] - V[3][2]}};
  const double b[3] = {-V[3][0], -V[3][1], -V[3][2]};
  double inv[3][3];
  inv[0][0] =   m[1][1]*m[2][2] - m[1][2]*m[2][1];
  inv[0][1] = - m[0][1]*m[2][2] + m[0][2]*m[2][1];
  inv[0][2] =   m[0][1]*m[1][2] - m[0][2]*m[1][1];
  inv[1][0] = - m[1][0]*m[2][2] + m[1][2]*m[2][0];
  inv[1][1] =   m[0][0]*m[2][2] - m[0][2]*m[2][0];
  inv[1][2] = - m[0][0]*m[1][2] + m[0][2]*m[1][0];
  inv[2][0] =   m[1][0]*m[2][1] - m[1][1]*m[2][0];
  inv[2][1] = - m[0][0]*m[2][1] + m[0][1]*m[2][0];
  inv[2][2] =   m[0][0]*m[1][1] - m[0][1]*m[1][0];
  const double det = m[0][0]*inv[0][0] + m[0][1]*inv[1][0] + m[0][2]*inv[2][0];
  const double invdet = 1.0 / det;
  for (int i = 0; i < 3; i ++) for (int j = 0; j < 3; j ++) inv[i][j] = inv[i][j] * invdet;
  l[0] = inv[0][0] * b[0] + inv[0][1] * b[1] + inv[0][2] * b[2];
  l[1] = inv[1][0] * b[0] + inv[1][1] * b[1] + inv[1][2] * b[2];
  l[2] = inv[2][0] * b[0] + inv[2][1] * b[1] + inv[2][2] * b[2];
  l[3] = 1.0 - l[0] - l[1] - l[2];
  return l[0] >= -eps && l[0] < 1.0 + eps && l[1] >= -eps && l[1] < 1.0 + eps &&
         l[2] >= -eps && l[2] < 1.0 + eps && l[3] >= -eps && l[3] < 1.0 + eps;
}

/* ref: include/ftk/numeric/critical_point_type.hh:40-72; eigen_solver2.hh:18-41,63-68;
 *      characteristic_polynomial.hh:10-17; quadratic_solver.hh:12-25 */
uint32_t cpo_cp_type_2d(const double J[2][2], int symmetric)
{
  if (symmetric) {
    const double m00 = J[0][0], m10 = J[1][0], m11 = J[1][1];
    const double b = -(m00 + m11), c = m00 * m11 - m10 * m10;
    const double delta = fma(b, b, -4 * c);
    const double sqrt_delta = delta < 0 ? 0 : sqrt(delta);
    double e0 = 0.5 * (-b + sqrt_delta), e1 = 0.5 * (-b - sqrt_delta);
    if (fabs(e0) < fabs(e1)) { const double t = e0; e0 = e1; e1 = t; }
    if (e0 > 0 && e1 > 0) return 2;
    else if (e0 < 0 && e1 < 0) return 8;
    else if (e0 * e1 < 0) return 4;
    else return 1;
  } else {
    const double P2 = 1.0, P1 = -(J[0][0] + J[1][1]), P0 = J[0][0] * J[1][1] - J[1][0] * J[0][1];
    const double delta = P1 * P1 - 4 * P2 * P0;
    if (delta >= 0) {
      const double r0 = (-P1 + sqrt(delta)) / (2 * P2), r1 = (-P1 - sqrt(delta)) / (2 * P2);
      if (r0 * r1 < 0) return 4;
      else if (r0 > 0 && r1 > 0) return 2;
      else if (r0 < 0 && r1 < 0) return 8;
      else return 1;
    } else {
      /* conjugate roots (or NaN delta): x0 = (-P1 + complex_sqrt(delta)) / (2 P2) with
       * complex_sqrt(z) = std::pow(std::complex<T>(z), 0.5)  (ref: numeric/sqrt.hh:9-15).
       * libstdc++ evaluates that as polar(exp(0.5 * log|z|), 0.5 * arg(z)), so the real part of
       * the root is not exactly -P1/2: it carries rho * cos(pi/2) ~ rho * 6.1e-17, and a NaN
       * delta makes it NaN (=> "center"). */
      const double rho = exp(0.5 * log(fabs(delta))), theta = 0.5 * atan2(0.0, delta);
      const double re = (-P1 + rho * cos(theta)) / (2 * P2);
      if (re < 0) return 16;
      else if (re > 0) return 32;
      else return 64;
    }
  }
}

/* ref: critical_point_type.hh:74-93; eigen_solver3.hh:16-47; characteristic_polynomial.hh:36-47 */
uint32_t cpo_cp_type_3d(const double A[3][3], int symmetric)
{
  if (!symmetric) return 0;
  const double b = -(A[0][0] + A[1][1] + A[2][2]);
  const double c = A[1][1]*A[2][2] + A[0][0]*A[2][2] + A[0][0]*A[1][1]
    - A[0][1]*A[1][0] - A[1][2]*A[2][1] - A[0][2]*A[2][0];
  const double d = -(A[0][0] * (A[1][1]*A[2][2] - A[1][2]*A[2][1])
                   - A[0][1] * (A[1][0]*A[2][2] - A[1][2]*A[2][0])
                   + A[0][2] * (A[1][0]*A[2][1] - A[1][1]*A[2][0]));
  double x[3], disc, q, r, dum1, term1, r13;
  q = (3.0*c - (b*b))/9.0;
  r = (-(27.0*d) + b*(9.0*c - 2.0*(b*b)))/54.0;
  disc = q*q*q + r*r;
  term1 = (b/3.0);
  if (disc >= 0) {
    r13 = ((r < 0) ? -pow(-r,(1.0/3.0)) : pow(r,(1.0/3.0)));
    x[0] = -term1 + 2.0*r13;
    x[1] = -(r13 + term1);
    x[2] = -(r13 + term1);
  } else {
    q = -q;
    dum1 = q*q*q;
    dum1 = acos(r/sqrt(dum1));
    r13 = 2.0*sqrt(q);
    x[0] = -term1 + r13*cos(dum1/3.0);
    x[1] = -term1 + r13*cos((dum1 + 2.0*M_PI)/3.0);
    x[2] = -term1 + r13*cos((dum1 + 4.0*M_PI)/3.0);
  }
  if (x[0] * x[1] * x[2] == 0.0) return 1;
  if (x[0] < 0 && x[1] < 0 && x[2] < 0) return 8;
  else if (x[0] > 0 && x[1] > 0 && x[2] > 0) return 2;
  else return 4;
}

/* ------------------------------------------------------------------------------------------
 * 4. field derivation     ref: include/ftk/ndarray/grad.hh
 * ---------------------------------------------------------------------------------------- */
static int clampi(int i, int n) { return i < 0 ? 0 : (i > n - 1 ? n - 1 : i); }

/* ref: grad.hh:10-31 gradient2D */
void cpo_gradient2D(const double *s, int DW, int DH, double *g)
{
#define F(i, j) s[(size_t)clampi(i, DW) + (size_t)DW * clampi(j, DH)]
  for (int j = 0; j < DH; j ++)
    for (int i = 0; i < DW; i ++) {
      g[0 + 2 * ((size_t)i + (size_t)DW * j)] = (F(i+1, j) - F(i-1, j)) * (DW-1);
      g[1 + 2 * ((size_t)i + (size_t)DW * j)] = (F(i, j+1) - F(i, j-1)) * (DH-1);
    }
#undef F
}

/* ref: grad.hh:54-86 jacobian2D<T, symmetric> -- including the operator-precedence quirk
 * (only the second term is scaled) and the two-index writes of the non-symmetric branch,
 * which land on elements (0,1,0,0) and (1,0,0,0). */
void cpo_jacobian2D(const double *v, int DW, int DH, int symmetric, double *J)
{
  memset(J, 0, sizeof(double) * 4 * (size_t)DW * DH);
#define F(c, i, j) v[(c) + 2 * ((size_t)clampi(i, DW) + (size_t)DW * clampi(j, DH))]
#define G(a, b, i, j) J[(a) + 2 * ((b) + 2 * ((size_t)(i) + (size_t)DW * (j)))]
  for (int j = 0; j < DH; j ++)
    for (int i = 0; i < DW; i ++) {
      const double H00 = F(0, i+1, j) - F(0, i-1, j) * (DW-1),
                   H01 = F(0, i, j+1) - F(0, i, j-1) * (DH-1),
                   H10 = F(1, i+1, j) - F(1, i-1, j) * (DW-1),
                   H11 = F(1, i, j+1) - F(1, i, j-1) * (DH-1);
      G(0, 0, i, j) = H00;
      G(1, 1, i, j) = H11;
      if (symmetric) G(0, 1, i, j) = G(1, 0, i, j) = (H01 + H10) * 0.5;
      else { J[0 + 1 * 2] = H01; J[1 + 0 * 2] = H10; } /* grad(0,1) = H01; grad(1,0) = H10 */
    }
#undef F
#undef G
}

/* ref: grad.hh:130-149 gradient3D (interior only; border stays zero) */
void cpo_gradient3D(const double *s, int DW, int DH, int DD, double *g)
{
  memset(g, 0, sizeof(double) * 3 * (size_t)DW * DH * DD);
#define S(i, j, k) s[(size_t)(i) + (size_t)DW * ((size_t)(j) + (size_t)DH * (k))]
#define G(c, i, j, k) g[(c) + 3 * ((size_t)(i) + (size_t)DW * ((size_t)(j) + (size_t)DH * (k)))]
  for (int k = 1; k < DD-1; k ++)
    for (int j = 1; j < DH-1; j ++)
      for (int i = 1; i < DW-1; i ++) {
        G(0, i, j, k) = 0.5 * (S(i+1, j, k) - S(i-1, j, k));
        G(1, i, j, k) = 0.5 * (S(i, j+1, k) - S(i, j-1, k));
        G(2, i, j, k) = 0.5 * (S(i, j, k+1) - S(i, j, k-1));
      }
#undef S
#undef G
}

/* ref: grad.hh:175-212 jacobian3D (b = 2) */
void cpo_jacobian3D(const double *V, int DW, int DH, int DD, double *J)
{
  memset(J, 0, sizeof(double) * 9 * (size_t)DW * DH * DD);
#define VV(c, i, j, k) V[(c) + 3 * ((size_t)(i) + (size_t)DW * ((size_t)(j) + (size_t)DH * (k)))]
#define JJ(a, b, i, j, k) J[(a) + 3 * ((b) + 3 * ((size_t)(i) + (size_t)DW * ((size_t)(j) + (size_t)DH * (k))))]
  for (int k = 2; k < DD-2; k ++)
    for (int j = 2; j < DH-2; j ++)
      for (int i = 2; i < DW-2; i ++)
        for (int c = 0; c < 3; c ++) {
          JJ(c, 0, i, j, k) = 0.5 * (VV(c, i+1, j, k) - VV(c, i-1, j, k));
          JJ(c, 1, i, j, k) = 0.5 * (VV(c, i, j+1, k) - VV(c, i, j-1, k));
          JJ(c, 2, i, j, k) = 0.5 * (VV(c, i, j, k+1) - VV(c, i, j, k-1));
        }
#undef VV
#undef JJ
}

/* ref: include/ftk/ndarray.hh:769-779 ndarray<T>::resolution */
double cpo_array_resolution(const double *p, uint64_t n)
{
  double r = DBL_MAX;
  for (uint64_t i = 0; i < n; i ++) if (p[i] != 0.0) r = dmin(r, fabs(p[i]));
  return r;
}

/* ------------------------------------------------------------------------------------------
 * 5. the tracker          ref: include/ftk/filters/critical_point_tracker{,_2d_regular,_3d_regular}.hh
 * ---------------------------------------------------------------------------------------- */
typedef struct { double *scalar, *vector, *jacobian; } snapshot_t;

struct cpo_ctx {
  cpo_config cfg;
  const mesh_t *m;
  int n;                                    /* spatial dims = cfg.nd */
  size_t nvert;                             /* vertices per layer */
  snapshot_t snaps[4]; int nsnaps;          /* std::deque<field_data_snapshot_t> */
  int current_timestep;
  double resolution; uint64_t factor;       /* critical_point_tracker.hh:162-163 */
  cpo_point *pts; size_t npts, cap;         /* discrete_critical_points */
  int sorted;
  /* finalize results */
  uint64_t *labels; int32_t *deg;
  uint64_t ntraj; uint64_t *traj_off, *traj_idx; uint8_t *traj_loop;
  /* physical coordinates, regular_tracker.hh:12-17,38-40,49-52 */
  int coords_mode; double *coords; size_t ncoords;
};

/* set_coords_bounds (mode 1: 2*nd doubles) / set_coords_rectilinear (mode 2: x[W] y[H] [z[D]]) /
 * set_coords_explicit (mode 3: (ncomp, W, H), ncomp = n / (W*H));  ref: regular_tracker.hh:38-40 */
void cpo_set_coords(cpo_ctx *c, int mode, const double *data, uint64_t n)
{
  free(c->coords); c->coords = NULL;
  c->coords_mode = mode; c->ncoords = n;
  if (mode && n) { c->coords = (double *)malloc(sizeof(double) * n); memcpy(c->coords, data, sizeof(double) * n); }
}

/* simplex_coordinates: critical_point_tracker_2d_regular.hh:494-526, ..._3d_regular.hh:343-379.
 * vt = vertex (x, y[, z], t); the array domain starts at 0.  Note the 3D explicit mode as the reference has it:
 * coordinates looked up by (x, y) only and the time slot filled with z. */
static void simplex_coordinates(const cpo_ctx *c, const int *vt, double X[4])
{
  const int n = c->n;
  const int32_t *dims = c->cfg.dims;
  if (c->coords_mode == 1) {
    for (int j = 0; j < n; j ++)
      X[j] = ((vt[j] - 0) / (double)(dims[j] - 1)) * (c->coords[2*j+1] - c->coords[2*j]) + c->coords[2*j];
    if (n == 2) X[2] = 0.0;
    X[3] = vt[n];
  } else if (c->coords_mode == 2) {
    size_t off = 0;
    for (int j = 0; j < n; j ++) { X[j] = c->coords[off + vt[j]]; off += dims[j]; }
    if (n == 2) X[2] = 0.0;
    X[3] = vt[n];
  } else if (c->coords_mode == 3) {
    const size_t nc = c->ncoords / ((size_t)dims[0] * dims[1]);
    const size_t k = (size_t)vt[0] + (size_t)dims[0] * vt[1];
    X[0] = c->coords[0 + nc * k]; X[1] = c->coords[1 + nc * k];
    if (n == 2) X[2] = nc > 2 ? c->coords[2 + nc * k] : 0.0;
    else X[2] = c->coords[2 + nc * k];
    X[3] = vt[2];
  } else {
    for (int j = 0; j < n; j ++) X[j] = vt[j];
    if (n == 2) X[2] = 0.0;
    X[3] = vt[n];
  }
}

cpo_ctx *cpo_create(const cpo_config *cfg)
{
  if (cfg->nd != 2 && cfg->nd != 3) return NULL;
  cpo_ctx *c = (cpo_ctx *)calloc(1, sizeof(cpo_ctx));
  c->cfg = *cfg;
  c->n = cfg->nd;
  c->m = get_mesh(cfg->nd + 1);
  c->nvert = (size_t)cfg->dims[0] * cfg->dims[1] * (cfg->nd == 3 ? cfg->dims[2] : 1);
  c->current_timestep = cfg->start_timestep;
  c->resolution = DBL_MAX; c->factor = 1;
  return c;
}

static void free_snapshot(snapshot_t *s) { free(s->scalar); free(s->vector); free(s->jacobian); memset(s, 0, sizeof(*s)); }

void cpo_destroy(cpo_ctx *c)
{
  if (c) free(c->coords);
  if (!c) return;
  for (int i = 0; i < c->nsnaps; i ++) free_snapshot(&c->snaps[i]);
  free(c->pts); free(c->labels); free(c->deg); free(c->traj_off); free(c->traj_idx); free(c->traj_loop);
  free(c);
}

static double *dup_array(const double *p, size_t n)
{
  double *q = (double *)malloc(sizeof(double) * n);
  memcpy(q, p, sizeof(double) * n);
  return q;
}

/* ref: critical_point_tracker_2d_regular.hh:238-261, critical_point_tracker_3d_regular.hh:125-148,
 *      critical_point_tracker.hh:202-213 */
int cpo_push_snapshot(cpo_ctx *c, const double *scalar, const double *vector, const double *jacobian)
{
  if (c->nsnaps >= 4) return -1;
  const int n = c->n; const int *d = c->cfg.dims;
  snapshot_t s; memset(&s, 0, sizeof(s));
  if (scalar) s.scalar = dup_array(scalar, c->nvert);
  if (vector) s.vector = dup_array(vector, c->nvert * n);
  else if (c->cfg.vector_source == CPO_SOURCE_DERIVED && scalar) {
    s.vector = (double *)malloc(sizeof(double) * c->nvert * n);
    if (n == 2) cpo_gradient2D(scalar, d[0], d[1], s.vector); else cpo_gradient3D(scalar, d[0], d[1], d[2], s.vector);
  }
  if (jacobian) s.jacobian = dup_array(jacobian, c->nvert * n * n);
  else if (c->cfg.jacobian_source == CPO_SOURCE_DERIVED && s.vector) {
    s.jacobian = (double *)malloc(sizeof(double) * c->nvert * n * n);
    /* 2D scalar path: jacobian2D<double, true>; 2D vector path: jacobian2D<double, false> */
    if (n == 2) cpo_jacobian2D(s.vector, d[0], d[1], vector ? 0 : 1, s.jacobian);
    else cpo_jacobian3D(s.vector, d[0], d[1], d[2], s.jacobian);
  }
  c->snaps[c->nsnaps ++] = s;
  return 0;
}

/* ref: critical_point_tracker.hh:850-864 update_vector_field_scaling_factor(minbits=8, maxbits=21) */
static void update_scaling_factor(cpo_ctx *c)
{
  for (int i = 0; i < c->nsnaps; i ++)
    if (c->snaps[i].vector)
      c->resolution = dmin(c->resolution, cpo_array_resolution(c->snaps[i].vector, c->nvert * c->n));
  int nbits = (int)ceil(log2(1.0 / c->resolution));
  const int minbits = 8, maxbits = 21;
  nbits = nbits < maxbits ? nbits : maxbits;  /* std::min(nbits, maxbits) */
  nbits = minbits > nbits ? minbits : nbits;  /* std::max(minbits, ...) */
  c->factor = (uint64_t)(1 << nbits);
}

double cpo_scaling_factor(const cpo_ctx *c) { return (double)c->factor; }
double cpo_resolution(const cpo_ctx *c) { return c->resolution; }

typedef struct { cpo_point *p; size_t n, cap; } ptvec_t;
static void ptvec_push(ptvec_t *v, const cpo_point *p)
{
  if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 256; v->p = (cpo_point *)realloc(v->p, sizeof(cpo_point) * v->cap); }
  v->p[v->n ++] = *p;
}

/* array index of a vertex: ref ndarray.hh:129-135 (dim 0 fastest); local_array_domain.start == 0 */
static size_t vidx(const cpo_ctx *c, const int *v)
{
  const int *d = c->cfg.dims;
  return c->n == 2 ? (size_t)v[0] + (size_t)d[0] * v[1]
                   : (size_t)v[0] + (size_t)d[0] * ((size_t)v[1] + (size_t)d[1] * v[2]);
}

/* ref: regular_tracker.hh:188-194 simplex_indices + lattice.hh:196-207 to_integer on the mesh
 * lattice (starts = domain lb + time 0; prod = 1, W', W'H'[, W'H'D']), uint64 truncated to int */
static int32_t sos_index(const cpo_ctx *c, const int *v /* nd+1 coords */)
{
  uint64_t prod = 1, i = 0;
  for (int j = 0; j <= c->n; j ++) {
    const int lbj = j < c->n ? c->cfg.lb[j] : 0;
    const int idx = v[j] - lbj;
    if (j == 0) i = (uint64_t)(int64_t)idx;          /* uint i(idx[0]) */
    else i += (uint64_t)((int64_t)idx) * prod;        /* idx[j] * prod_[j] (int -> size_t) */
    if (j < c->n) prod *= (uint64_t)(c->cfg.ub[j] - c->cfg.lb[j] + 1);
  }
  return (int32_t)(uint32_t)i;
}

/* ref: simplicial_regular_mesh.hh:356-371 valid(): every vertex coordinate within [lb, ub];
 * time bounds are [0, INT_MAX] (tracker.hh:75-76, regular_tracker.hh:117-123) */
static int simplex_valid(const cpo_ctx *c, int verts[][MAXND], int nv)
{
  for (int i = 0; i < nv; i ++) {
    for (int j = 0; j < c->n; j ++)
      if (verts[i][j] < c->cfg.lb[j] || verts[i][j] > c->cfg.ub[j]) return 0;
    if (verts[i][c->n] < 0) return 0;
  }
  return 1;
}

/* ref: critical_point_tracker_2d_regular.hh:584-685 check_simplex */
static int check_simplex_2d(const cpo_ctx *c, const int corner[3], int type, cpo_point *cp)
{
  const mesh_t *m = c->m;
  int vt[3][MAXND];
  for (int i = 0; i < 3; i ++) for (int j = 0; j < 3; j ++) vt[i][j] = corner[j] + m->unit[2][type].v[i][j];
  if (!simplex_valid(c, vt, 3)) return 0;

  double v[3][2]; const snapshot_t *sn[3];
  for (int i = 0; i < 3; i ++) {
    sn[i] = &c->snaps[vt[i][2] == c->current_timestep ? 0 : 1];
    const size_t k = vidx(c, vt[i]);
    v[i][0] = sn[i]->vector[2 * k]; v[i][1] = sn[i]->vector[2 * k + 1];
  }
  i64 vf[3][2];
  for (int i = 0; i < 3; i ++)
    for (int j = 0; j < 2; j ++) {
      const double x = v[i][j];
      if (isnan(x) || isinf(x)) return 0;
      vf[i][j] = (i64)(x * (double)c->factor);   /* v * uint64 factor -> double product -> int64 */
    }
  int32_t indices[3];
  for (int i = 0; i < 3; i ++) indices[i] = sos_index(c, vt[i]);
  if (!cpo_robust_cp_in_simplex2(vf, indices)) return 0;

  double mu[3];
  const int succ2 = inverse_lerp_s2v2(v, mu);
  if (!succ2) clamp_barycentric(3, mu);

  /* simplex_coordinates (REGULAR_COORDS_SIMPLE) + lerp_s2v4: ref linear_interpolation.hh:81-89 */
  double X[3][4];
  for (int i = 0; i < 3; i ++) simplex_coordinates(c, vt[i], X[i]);
  double x[4];
  for (int k = 0; k < 4; k ++) x[k] = X[0][k] * mu[0] + X[1][k] * mu[1] + X[2][k] * mu[2];
  memset(cp, 0, sizeof(*cp));
  cp->x[0] = x[0]; cp->x[1] = x[1]; cp->x[2] = x[2]; cp->t = x[3];

  if (c->cfg.scalar_source != CPO_SOURCE_NONE) {
    double values[3];
    for (int i = 0; i < 3; i ++) values[i] = sn[i]->scalar[vidx(c, vt[i])];
    cp->scalar = values[0] * mu[0] + values[1] * mu[1] + values[2] * mu[2];   /* lerp_s2 */
  }
  cp->corner[0] = corner[0]; cp->corner[1] = corner[1]; cp->corner[2] = 0; cp->corner[3] = corner[2];
  cp->simplex_type = type;
  cp->ordinal = m->is_ord[2][type];
  cp->timestep = c->current_timestep;

  if (c->cfg.compute_degrees) {
    if (cp->ordinal) {
      int deg = cpo_positive2(vf, indices);
      const int chi = type == 4 ? 1 : -1;
      deg *= chi;
      cp->cp_type = deg == 1 ? 1 : 2;
    } else cp->cp_type = 0;
  } else {
    double J[2][2] = {{0, 0}, {0, 0}};
    if (c->cfg.jacobian_source != CPO_SOURCE_NONE) {
      double Js[3][2][2];
      for (int i = 0; i < 3; i ++) {
        const size_t k = vidx(c, vt[i]);
        for (int a = 0; a < 2; a ++) for (int b = 0; b < 2; b ++)
          Js[i][a][b] = sn[i]->jacobian[b + 2 * (a + 2 * k)];        /* jacobian(k=b, j=a, x, y) */
      }
      for (int a = 0; a < 2; a ++) for (int b = 0; b < 2; b ++)       /* lerp_s2m2x2 */
        J[a][b] = Js[0][a][b] * mu[0] + Js[1][a][b] * mu[1] + Js[2][a][b] * mu[2];
      const double sym = 0.5 * (J[0][1] + J[1][0]);                   /* make_symmetric2x2 */
      J[0][1] = J[1][0] = sym;
    }
    cp->cp_type = cpo_cp_type_2d(J, c->cfg.jacobian_symmetric);
  }
  /* type filter: ref critical_point_tracker_2d_regular.hh:280, critical_point_tracker.hh:190-200 */
  if (c->cfg.use_type_filter && !(c->cfg.type_filter & cp->cp_type)) return 0;
  return 1;
}

/* ref: critical_point_tracker_3d_regular.hh:425-514 check_simplex */
static int check_simplex_3d(const cpo_ctx *c, const int corner[4], int type, cpo_point *cp)
{
  const mesh_t *m = c->m;
  int vt[4][MAXND];
  for (int i = 0; i < 4; i ++) for (int j = 0; j < 4; j ++) vt[i][j] = corner[j] + m->unit[3][type].v[i][j];
  if (!simplex_valid(c, vt, 4)) return 0;

  double v[4][3]; const snapshot_t *sn[4];
  for (int i = 0; i < 4; i ++) {
    sn[i] = &c->snaps[vt[i][3] == c->current_timestep ? 0 : 1];
    const size_t k = vidx(c, vt[i]);
    for (int j = 0; j < 3; j ++) v[i][j] = sn[i]->vector[3 * k + j];
  }
  double mu[4];
  const int succ2 = inverse_lerp_s3v3(v, mu);

  if (c->cfg.robust_detection) {
    i64 vf[4][3];
    for (int i = 0; i < 4; i ++)
      for (int j = 0; j < 3; j ++) {
        const double x = v[i][j];
        if (isnan(x) || isinf(x)) return 0;
        vf[i][j] = (i64)(x * (double)c->factor);
      }
    int32_t indices[4];
    for (int i = 0; i < 4; i ++) indices[i] = sos_index(c, vt[i]);
    if (!cpo_robust_cp_in_simplex3(vf, indices)) return 0;
  } else {
    if (!succ2) return 0;
  }
  clamp_barycentric(4, mu);

  double X[4][4], x[4];
  for (int i = 0; i < 4; i ++) simplex_coordinates(c, vt[i], X[i]);
  for (int k = 0; k < 4; k ++) /* lerp_s3v4 */
    x[k] = X[0][k] * mu[0] + X[1][k] * mu[1] + X[2][k] * mu[2] + X[3][k] * mu[3];
  memset(cp, 0, sizeof(*cp));
  cp->x[0] = x[0]; cp->x[1] = x[1]; cp->x[2] = x[2]; cp->t = x[3];

  if (c->cfg.scalar_source != CPO_SOURCE_NONE) {
    double values[4];
    for (int i = 0; i < 4; i ++) values[i] = sn[i]->scalar[vidx(c, vt[i])];
    cp->scalar = values[0] * mu[0] + values[1] * mu[1] + values[2] * mu[2] + values[3] * mu[3]; /* lerp_s3 */
  }
  double J[3][3];
  for (int a = 0; a < 3; a ++)
    for (int b = 0; b < 3; b ++) {  /* lerp_s3m3x3: v = 0; v += V[i] * mu[i] */
      double acc = 0.0;
      for (int i = 0; i < 4; i ++) {
        const double Jv = sn[i]->jacobian ? sn[i]->jacobian[b + 3 * (a + 3 * vidx(c, vt[i]))] : 0.0;
        acc += Jv * mu[i];
      }
      J[a][b] = acc;
    }
  cp->cp_type = cpo_cp_type_3d(J, c->cfg.jacobian_symmetric);
  for (int j = 0; j < 4; j ++) cp->corner[j] = corner[j];
  cp->simplex_type = type;
  cp->ordinal = m->is_ord[3][type];
  cp->timestep = c->current_timestep;
  return 1;
}

/* ref: simplicial_regular_mesh.hh:1030-1045 element_for + :480-493 from_work_index +
 *      lattice.hh:209-223 from_integer (x fastest); regular_tracker.hh:196-211 */
static void sweep(cpo_ctx *c, int ordinal)
{
  const mesh_t *m = c->m;
  const int n = c->n, d = n; /* n-simplices of the (n+1)-D mesh */
  const int nty = ordinal ? m->n_ord[d] : m->n_int[d];
  const int *types = ordinal ? m->ord_types[d] : m->int_types[d];
  size_t sz[3] = {1, 1, 1};
  for (int j = 0; j < n; j ++) sz[j] = (size_t)(c->cfg.ub[j] - c->cfg.lb[j] + 1);
  const int64_t ncorners = (int64_t)(sz[0] * sz[1] * sz[2]);
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = c->cfg.nthreads > 0 ? c->cfg.nthreads : omp_get_max_threads();
#endif
  ptvec_t *found = (ptvec_t *)calloc((size_t)nthreads, sizeof(ptvec_t));
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
  for (int64_t ci = 0; ci < ncorners; ci ++) {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    int corner[4];
    int64_t r = ci;
    corner[0] = (int)(r % (int64_t)sz[0]) + c->cfg.lb[0]; r /= (int64_t)sz[0];
    corner[1] = (int)(r % (int64_t)sz[1]) + c->cfg.lb[1]; r /= (int64_t)sz[1];
    if (n == 3) { corner[2] = (int)r + c->cfg.lb[2]; corner[3] = c->current_timestep; }
    else corner[2] = c->current_timestep;
    for (int it = 0; it < nty; it ++) {
      cpo_point cp;
      const int ok = n == 2 ? check_simplex_2d(c, corner, types[it], &cp) : check_simplex_3d(c, corner, types[it], &cp);
      if (ok) ptvec_push(&found[tid], &cp);
    }
  }
  for (int t = 0; t < nthreads; t ++) {
    for (size_t i = 0; i < found[t].n; i ++) {
      if (c->npts == c->cap) { c->cap = c->cap ? c->cap * 2 : 1024; c->pts = (cpo_point *)realloc(c->pts, sizeof(cpo_point) * c->cap); }
      c->pts[c->npts ++] = found[t].p[i];
    }
    free(found[t].p);
  }
  free(found);
  c->sorted = 0;
}

/* ref: critical_point_tracker_2d_regular.hh:263-433 / 3d:150-308 update_timestep (xl == NONE) */
int cpo_update_timestep(cpo_ctx *c)
{
  if (c->nsnaps < 1) return -1;
  update_scaling_factor(c);   /* 2D: non-GMP build only; 3D: always */
  sweep(c, 1);
  if (c->nsnaps >= 2) sweep(c, 0);
  return 0;
}

/* ref: critical_point_tracker.hh:841-848 advance_timestep, :231-237 pop_field_data_snapshot */
int cpo_advance_timestep(cpo_ctx *c)
{
  const int rc = cpo_update_timestep(c);
  if (c->nsnaps > 0) {
    free_snapshot(&c->snaps[0]);
    for (int i = 1; i < c->nsnaps; i ++) c->snaps[i-1] = c->snaps[i];
    c->nsnaps --;
    memset(&c->snaps[c->nsnaps], 0, sizeof(snapshot_t));
  }
  c->current_timestep ++;
  return rc;
}

/* element order: ref simplicial_regular_mesh.hh:327-337 (corner lexicographic x first, then type) */
static int cmp_point(const void *a_, const void *b_)
{
  const cpo_point *a = (const cpo_point *)a_, *b = (const cpo_point *)b_;
  for (int j = 0; j < 4; j ++) if (a->corner[j] != b->corner[j]) return a->corner[j] < b->corner[j] ? -1 : 1;
  return (a->simplex_type > b->simplex_type) - (a->simplex_type < b->simplex_type);
}

static void sort_points(cpo_ctx *c)
{
  if (c->sorted) return;
  qsort(c->pts, c->npts, sizeof(cpo_point), cmp_point);
  /* std::map assignment semantics: a later insert with the same key overwrites (keep one) */
  size_t w = 0;
  for (size_t i = 0; i < c->npts; i ++) {
    if (w > 0 && cmp_point(&c->pts[w-1], &c->pts[i]) == 0) c->pts[w-1] = c->pts[i];
    else c->pts[w ++] = c->pts[i];
  }
  c->npts = w;
  c->sorted = 1;
}

/* time-slab helpers (test infrastructure for the multi-GPU protocol; not reference functions): a slab
 * inherits the running minimum of critical_point_tracker.hh:850-864 from the slabs before it, and the
 * root merges the punctured simplices of every slab before trace_critical_points_offline */
void cpo_set_resolution(cpo_ctx *c, double r) { if (r > 0 && r < c->resolution) c->resolution = r; }
void cpo_import_points(cpo_ctx *c, const cpo_point *pts, uint64_t n)
{
  for (uint64_t i = 0; i < n; i ++) {
    if (c->npts == c->cap) { c->cap = c->cap ? c->cap * 2 : 1024; c->pts = (cpo_point *)realloc(c->pts, sizeof(cpo_point) * c->cap); }
    c->pts[c->npts ++] = pts[i];
  }
  c->sorted = 0;
}

uint64_t cpo_num_points(const cpo_ctx *c) { sort_points((cpo_ctx *)c); return c->npts; }
void cpo_get_points(const cpo_ctx *c, cpo_point *out) { sort_points((cpo_ctx *)c); memcpy(out, c->pts, sizeof(cpo_point) * c->npts); }

static int64_t find_point(const cpo_ctx *c, const int corner[4], int type)
{
  cpo_point key; memcpy(key.corner, corner, sizeof(int) * 4); key.simplex_type = type;
  int64_t lo = 0, hi = (int64_t)c->npts - 1;
  while (lo <= hi) {
    const int64_t mid = (lo + hi) / 2;
    const int r = cmp_point(&c->pts[mid], &key);
    if (r == 0) return mid;
    if (r < 0) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

/* neighbours(f) = union over c in side_of(f) of sides(c), restricted to punctured elements,
 * returned as sorted unique indices (including f itself).
 * ref: critical_point_tracker_2d_regular.hh:189-197, simplicial_regular_mesh.hh:571-601 */
static int hit_neighbors(const cpo_ctx *c, size_t i, int64_t *out)
{
  const mesh_t *m = c->m; const int n = c->n, nd = n + 1;
  const cpo_point *p = &c->pts[i];
  int corner[4];
  if (n == 2) { corner[0] = p->corner[0]; corner[1] = p->corner[1]; corner[2] = p->corner[3]; corner[3] = 0; }
  else memcpy(corner, p->corner, sizeof(corner));
  int cnt = 0;
  for (int a = 0; a < m->nsideof[n][p->simplex_type]; a ++) {
    const tyoff_t *cell = &m->sideof[n][p->simplex_type][a];
    int cc[4] = {0, 0, 0, 0};
    for (int j = 0; j < nd; j ++) cc[j] = corner[j] + cell->off[j];
    for (int b = 0; b < m->nsides[nd][cell->type]; b ++) {
      const tyoff_t *sd = &m->sides[nd][cell->type][b];
      int sc[4] = {0, 0, 0, 0};
      for (int j = 0; j < nd; j ++) sc[j] = cc[j] + sd->off[j];
      int key[4];
      if (n == 2) { key[0] = sc[0]; key[1] = sc[1]; key[2] = 0; key[3] = sc[2]; } else memcpy(key, sc, sizeof(key));
      const int64_t k = find_point(c, key, sd->type);
      if (k < 0) continue;
      int dup = 0; for (int q = 0; q < cnt; q ++) if (out[q] == k) dup = 1;
      if (!dup) out[cnt ++] = k;
    }
  }
  for (int a = 1; a < cnt; a ++) for (int b = a; b > 0 && out[b-1] > out[b]; b --) { const int64_t t = out[b]; out[b] = out[b-1]; out[b-1] = t; }
  return cnt;
}

/* ref: include/ftk/basic/duf.hh:41-72 (unite: hook the larger root under the smaller; find: path halving) */
static uint64_t uf_find(uint64_t *parent, uint64_t i)
{
  while (i != parent[i]) { parent[i] = parent[parent[i]]; i = parent[i]; }
  return i;
}
static void uf_unite(uint64_t *parent, uint64_t i, uint64_t j)
{
  i = uf_find(parent, i); j = uf_find(parent, j);
  if (i == j) return;
  if (i > j) { const uint64_t t = i; i = j; j = t; }
  parent[j] = i;
}

static int in_sorted(const int64_t *set, int n, int64_t v) { for (int i = 0; i < n; i ++) if (set[i] == v) return 1; return 0; }

/* ref: critical_point_tracker.hh:668-817 trace_critical_points_offline
 *      include/ftk/geometry/cc2curves.hh:10-122 connected_component_to_linear_components, is_loop
 *      include/ftk/algorithms/cca.hh:91-116 */
int cpo_finalize(cpo_ctx *c)
{
  sort_points(c);
  const size_t N = c->npts;
  free(c->labels); free(c->deg); free(c->traj_off); free(c->traj_idx); free(c->traj_loop);
  c->labels = (uint64_t *)malloc(sizeof(uint64_t) * (N + 1));
  c->deg = (int32_t *)malloc(sizeof(int32_t) * (N + 1));
  c->traj_off = (uint64_t *)calloc(N + 2, sizeof(uint64_t));
  c->traj_idx = (uint64_t *)malloc(sizeof(uint64_t) * (N + 1));
  c->traj_loop = (uint8_t *)calloc(N + 1, 1);
  c->ntraj = 0;

  /* neighbour lists among punctured elements */
  int64_t (*nb)[12] = (int64_t (*)[12])malloc(sizeof(int64_t[12]) * (N + 1));
  int *nnb = (int *)malloc(sizeof(int) * (N + 1));
  for (size_t i = 0; i < N; i ++) nnb[i] = hit_neighbors(c, i, nb[i]);

  /* (1) all components, special nodes included */
  uint64_t *parent = (uint64_t *)malloc(sizeof(uint64_t) * (N + 1));
  for (size_t i = 0; i < N; i ++) parent[i] = i;
  for (size_t i = 0; i < N; i ++) for (int a = 0; a < nnb[i]; a ++) uf_unite(parent, i, (uint64_t)nb[i][a]);
  for (size_t i = 0; i < N; i ++) c->labels[i] = uf_find(parent, i);

  /* (2) special nodes: more than two punctured neighbours other than itself (cc2curves.hh:19-31) */
  uint8_t *special = (uint8_t *)calloc(N + 1, 1);
  for (size_t i = 0; i < N; i ++) {
    int d = 0; for (int a = 0; a < nnb[i]; a ++) if ((size_t)nb[i][a] != i) d ++;
    c->deg[i] = d;
    special[i] = d > 2;
  }
  /* (3) components of ordinary nodes (cc2curves.hh:33-43) */
  for (size_t i = 0; i < N; i ++) parent[i] = i;
  for (size_t i = 0; i < N; i ++) {
    if (special[i]) continue;
    for (int a = 0; a < nnb[i]; a ++) if (!special[nb[i][a]]) uf_unite(parent, i, (uint64_t)nb[i][a]);
  }
  uint64_t *olab = (uint64_t *)malloc(sizeof(uint64_t) * (N + 1));
  for (size_t i = 0; i < N; i ++) olab[i] = special[i] ? UINT64_MAX : uf_find(parent, i);

  /* (4) order every component by the reference's walk (cc2curves.hh:46-108).  The roots of the
   * min-hooking union-find are the smallest members, so visiting seeds in ascending index order
   * visits components by ascending smallest element. */
  uint8_t *visited = (uint8_t *)calloc(N + 1, 1);
  int64_t *fwd = (int64_t *)malloc(sizeof(int64_t) * (N + 1)), *bwd = (int64_t *)malloc(sizeof(int64_t) * (N + 1));
  uint64_t pos = 0;
  for (size_t seed = 0; seed < N; seed ++) {
    if (special[seed] || olab[seed] != seed) continue;   /* seed = *c.begin() */
    const uint64_t lab = olab[seed];
    size_t nf = 0, nbk = 0;
    visited[seed] = 1;
    int64_t sn[12]; int nsn = 0;                          /* seed_neighbors: ordinary, != seed */
    for (int a = 0; a < nnb[seed]; a ++) if ((size_t)nb[seed][a] != seed && !special[nb[seed][a]]) sn[nsn ++] = nb[seed][a];
    for (int dir = 0; dir < 2; dir ++) {
      if (nsn == 0) break;
      int64_t current = dir == 0 ? sn[0] : sn[nsn - 1];
      while (1) {
        if (!visited[current]) {
          if (dir == 0) fwd[nf ++] = current; else bwd[nbk ++] = current;
          visited[current] = 1;
        }
        int found_next = 0;
        for (int a = 0; a < nnb[current]; a ++) {
          const int64_t q = nb[current][a];
          if (q != current && !special[q] && olab[q] == lab && !visited[q]) { found_next = 1; current = q; break; }
        }
        if (!found_next) break;
      }
      if (nsn == 1) break;
    }
    /* trace = reverse(bwd) + seed + fwd */
    const uint64_t start = pos;
    for (size_t k = nbk; k > 0; k --) c->traj_idx[pos ++] = (uint64_t)bwd[k-1];
    c->traj_idx[pos ++] = seed;
    for (size_t k = 0; k < nf; k ++) c->traj_idx[pos ++] = (uint64_t)fwd[k];
    /* is_loop: size > 1 and back in neighbors(front)  (cc2curves.hh:113-122) */
    const uint64_t len = pos - start;
    int loop = 0;
    if (len > 1) {
      const uint64_t front = c->traj_idx[start], back = c->traj_idx[pos - 1];
      loop = in_sorted(nb[front], nnb[front], (int64_t)back);
    }
    c->traj_loop[c->ntraj] = (uint8_t)loop;
    c->traj_off[c->ntraj + 1] = pos;
    c->ntraj ++;
  }
  free(visited); free(fwd); free(bwd); free(olab); free(special); free(parent); free(nb); free(nnb);
  return 0;
}

uint64_t cpo_num_trajectories(const cpo_ctx *c) { return c->ntraj; }
void cpo_get_trajectories(const cpo_ctx *c, uint64_t *offsets, uint64_t *idx, uint8_t *loop)
{
  memcpy(offsets, c->traj_off, sizeof(uint64_t) * (c->ntraj + 1));
  memcpy(idx, c->traj_idx, sizeof(uint64_t) * c->traj_off[c->ntraj]);
  memcpy(loop, c->traj_loop, c->ntraj);
}
void cpo_get_component_labels(const cpo_ctx *c, uint64_t *labels) { memcpy(labels, c->labels, sizeof(uint64_t) * c->npts); }
void cpo_get_degrees(const cpo_ctx *c, int32_t *deg) { memcpy(deg, c->deg, sizeof(int32_t) * c->npts); }

/* ------------------------------------------------------------------------------------------
 * 6. synthetic generators   ref: include/ftk/ndarray/synthetic.hh, include/ftk/ndarray/stream.hh
 * ---------------------------------------------------------------------------------------- */
/* ref: synthetic.hh:10-14, 32-48 synthetic_woven_2D(DW, DH, t, scaling_factor = 15) */
void cpo_gen_woven(int DW, int DH, double t, double *out)
{
  const double scaling_factor = 15;
  for (int j = 0; j < DH; j ++)
    for (int i = 0; i < DW; i ++) {
      const double x = (((double)i / (DW-1)) - 0.5) * scaling_factor,
                   y = (((double)j / (DH-1)) - 0.5) * scaling_factor;
      out[(size_t)i + (size_t)DW * j] = cos(x*cos(t)-y*sin(t))*sin(x*sin(t)+y*cos(t));
    }
}

/* ref: synthetic.hh:262-297 merger_function_2Dt, synthetic_merger_2D */
void cpo_gen_merger(int DW, int DH, double t, double *out)
{
  for (int j = 0; j < DH; j ++)
    for (int i = 0; i < DW; i ++) {
      double x = (((double)i / (DW-1)) - 0.5) * 4, y = (((double)j / (DH-1)) - 0.5) * 4;
      const double xp = x * cos(t) - y * sin(t), yp = x * sin(t) + y * cos(t);
      x = xp; y = yp;
      const double cx0 = sin(t - M_PI_2), cx1 = sin(t + M_PI_2), cy0 = 1e-4, cy1 = 1e-4;
      const double f0 = exp(-((x-cx0)*(x-cx0) + (y-cy0)*(y-cy0))), f1 = exp(-((x-cx1)*(x-cx1) + (y-cy1)*(y-cy1)));
      out[(size_t)i + (size_t)DW * j] = dmax(f0, f1);
    }
}

/* ref: synthetic.hh:332-354 synthetic_moving_extremum<T, N> */
void cpo_gen_moving_extremum(int nd, const int32_t *dims, const double *x0, const double *dir, double t, double *out)
{
  double xc[3];
  for (int j = 0; j < nd; j ++) xc[j] = x0[j] + dir[j] * t;
  const int W = dims[0], H = dims[1], D = nd == 3 ? dims[2] : 1;
  for (int k = 0; k < D; k ++)
    for (int j = 0; j < H; j ++)
      for (int i = 0; i < W; i ++) {
        const int xi[3] = {i, j, k};
        double d = 0;
        for (int q = 0; q < nd; q ++) d += pow(xi[q] - xc[q], 2.0);
        out[(size_t)i + (size_t)W * ((size_t)j + (size_t)H * k)] = d;
      }
}

/* ref: synthetic.hh:130-150 double_gyre, :193-217 synthetic_double_gyre (domain [0,2]x[0,1]) */
void cpo_gen_double_gyre(int DW, int DH, double time, double A, double omega, double epsilon, double *out)
{
  for (int j = 0; j < DH; j ++)
    for (int i = 0; i < DW; i ++) {
      const double x = ((double)i / (DW-1)) * 2, y = ((double)j / (DH-1));
      const double a = epsilon * sin(omega * time);
      const double b = 1 - 2 * epsilon * sin(omega * time);
      const double f = a * x * x + b * x;
      const double dfdx = 2 * a * x + b;
      const double u = -M_PI * A * sin(M_PI * f) * cos(M_PI * y);
      const double v =  M_PI * A * cos(M_PI * f) * sin(M_PI * y) * dfdx;
      out[0 + 2 * ((size_t)i + (size_t)DW * j)] = u;
      out[1 + 2 * ((size_t)i + (size_t)DW * j)] = v;
    }
}

/* ref: synthetic.hh:441-494 synthetic_tornado (vector (3, xs, ys, zs), component fastest) */
void cpo_gen_tornado(int xs, int ys, int zs, int time, double *out)
{
  const double SMALL = 0.00000000001;
  const double xdelta = 1.0 / (xs-1.0), ydelta = 1.0 / (ys-1.0), zdelta = 1.0 / (zs-1.0);
  for (int iz = 0; iz < zs; iz ++) {
    const double z = iz * zdelta;
    const double xc = 0.5 + 0.1*sin(0.04*time+10.0*z);
    const double yc = 0.5 + 0.1*cos(0.03*time+3.0*z);
    const double r = 0.1 + 0.4 * z*z + 0.1 * z * sin(8.0*z);
    const double r2 = 0.2 + 0.1*z;
    for (int iy = 0; iy < ys; iy ++) {
      const double y = iy * ydelta;
      for (int ix = 0; ix < xs; ix ++) {
        const double x = ix * xdelta;
        double temp = sqrt( (y-yc)*(y-yc) + (x-xc)*(x-xc) );
        double scale = fabs( r - temp );
        if ( scale > r2 ) scale = 0.8 - scale;
        else scale = 1.0;
        double z0 = 0.1 * (0.1 - temp*z );
        if ( z0 < 0.0 ) z0 = 0.0;
        temp = sqrt( temp*temp + z0*z0 );
        scale = (r + r2 - temp) * scale / (temp + SMALL);
        scale = scale / (1+z);
        *out++ = scale * (y-yc) + 0.1*(x-xc);
        *out++ = scale * -(x-xc) + 0.1*(y-yc);
        *out++ = scale * z0;
      }
    }
  }
}

/* ref: synthetic.hh:239-260 synthetic_abc_flow */
void cpo_gen_abc(int DW, int DH, int DD, double A, double B, double C, double *out)
{
  for (int k = 0; k < DD; k ++)
    for (int j = 0; j < DH; j ++)
      for (int i = 0; i < DW; i ++) {
        const double x = (((double)i / (DW-1))) * 2 * M_PI,
                     y = (((double)j / (DH-1))) * 2 * M_PI,
                     z = (((double)k / (DD-1))) * 2 * M_PI;
        double *o = out + 3 * ((size_t)i + (size_t)DW * ((size_t)j + (size_t)DH * k));
        o[0] = A * sin(z) + C * cos(y);
        o[1] = B * sin(x) + A * cos(z);
        o[2] = C * sin(y) + B * cos(x);
      }
}
