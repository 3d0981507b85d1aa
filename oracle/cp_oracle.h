/* TEST INFRASTRUCTURE ONLY -- this is the parity oracle, never the product path.
 *
 * cp_oracle: a plain-C, CPU restatement of the critical-point extraction/tracking path of
 * hguo/ftk (CPU tracker, non-GMP build).  Every function in cp_oracle.c cites the reference
 * file:line it follows (paths relative to the reference checkout).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this restatement against fixtures
 * produced by the unmodified reference itself (oracle/ref_harness.cpp -> tests/golden/).
 */
#ifndef CP_ORACLE_H
#define CP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { CPO_SOURCE_NONE = 0, CPO_SOURCE_GIVEN = 1, CPO_SOURCE_DERIVED = 2 };

typedef struct {
  int32_t nd;                 /* 2 or 3 (spatial dimensionality) */
  int32_t dims[3];            /* array dims W, H, D */
  int32_t lb[3], ub[3];       /* tracker domain, inclusive (regular_tracker.hh:24) */
  int32_t scalar_source, vector_source, jacobian_source;
  int32_t jacobian_symmetric;
  int32_t robust_detection;   /* 3D only (critical_point_tracker_3d_regular.hh:442) */
  int32_t compute_degrees;    /* 2D only */
  int32_t use_type_filter;    /* 2D only */
  uint32_t type_filter;
  int32_t start_timestep;     /* initial current_timestep (tracker.hh:76); mesh time lb stays 0 */
  int32_t nthreads;           /* OpenMP threads for the sweep; <=0: default */
} cpo_config;

typedef struct {
  int32_t corner[4];          /* x, y, z, t (z = 0 in 2D) */
  int32_t simplex_type;       /* type index among all n-simplex types of the (n+1)-D mesh */
  int32_t ordinal;
  int32_t timestep;
  uint32_t cp_type;
  double x[3], t, scalar;
} cpo_point;

typedef struct cpo_ctx cpo_ctx;

cpo_ctx *cpo_create(const cpo_config *cfg);
void cpo_destroy(cpo_ctx *);
/* any pointer may be NULL according to the *_source settings; arrays are dim-0-fastest
 * scalar (W,H[,D]); vector (n,W,H[,D]); jacobian (n,n,W,H[,D]) */
int cpo_push_snapshot(cpo_ctx *, const double *scalar, const double *vector, const double *jacobian);
/* physical coordinates (regular_tracker.hh:38-40): mode 0 grid units, 1 bounds, 2 rectilinear, 3 explicit */
void cpo_set_coords(cpo_ctx *, int mode, const double *data, uint64_t n);
int cpo_update_timestep(cpo_ctx *);
int cpo_advance_timestep(cpo_ctx *);
int cpo_finalize(cpo_ctx *);
double cpo_scaling_factor(const cpo_ctx *);
double cpo_resolution(const cpo_ctx *);

/* time-slab helpers (not reference functions; see cp_oracle.c) */
void cpo_set_resolution(cpo_ctx *, double r);
void cpo_import_points(cpo_ctx *, const cpo_point *pts, uint64_t n);

uint64_t cpo_num_points(const cpo_ctx *);
/* sorted by the reference's element order (corner lexicographic x first, then type) */
void cpo_get_points(const cpo_ctx *, cpo_point *out);
uint64_t cpo_num_trajectories(const cpo_ctx *);
/* CSR: offsets[ntraj+1]; idx[] index into the sorted point array; loop[ntraj] */
void cpo_get_trajectories(const cpo_ctx *, uint64_t *offsets, uint64_t *idx, uint8_t *loop);
/* label of the full connected component (special nodes included) for each sorted point:
 * the index of the smallest point of the component */
void cpo_get_component_labels(const cpo_ctx *, uint64_t *labels);
/* degree (number of punctured neighbours) of each sorted point */
void cpo_get_degrees(const cpo_ctx *, int32_t *deg);

/* mesh tables (simplicial_regular_mesh.hh:620-831): nd_mesh = 3 or 4 */
int cpo_mesh_ntypes(int nd_mesh, int k, int scope /*0 all 1 ordinal 2 interval*/);
/* vertices of unit k-simplex `type`: out[(k+1) * nd_mesh] offsets in {0,1} */
void cpo_mesh_unit_simplex(int nd_mesh, int k, int type, int32_t *out);
int cpo_mesh_scope_type(int nd_mesh, int k, int scope, int itype);
/* out[] = n entries of (type, off[nd_mesh]); returns n */
int cpo_mesh_sides(int nd_mesh, int k, int type, int32_t *out);
int cpo_mesh_side_of(int nd_mesh, int k, int type, int32_t *out);

/* predicates (sign_det.hh / critical_point_test.hh) */
int cpo_robust_cp_in_simplex2(const int64_t V[3][2], const int32_t idx[3]);
int cpo_robust_cp_in_simplex3(const int64_t V[4][3], const int32_t idx[4]);
int cpo_positive2(const int64_t V[3][2], const int32_t idx[3]);
int cpo_positive3(const int64_t V[4][3], const int32_t idx[4]);

/* derivation (grad.hh) */
void cpo_gradient2D(const double *s, int W, int H, double *out);
void cpo_jacobian2D(const double *v, int W, int H, int symmetric, double *out);
void cpo_gradient3D(const double *s, int W, int H, int D, double *out);
void cpo_jacobian3D(const double *v, int W, int H, int D, double *out);
double cpo_array_resolution(const double *p, uint64_t n);

/* type classification (critical_point_type.hh) */
uint32_t cpo_cp_type_2d(const double J[2][2], int symmetric);
uint32_t cpo_cp_type_3d(const double J[3][3], int symmetric);

/* synthetic generators (ndarray/synthetic.hh, ndarray/stream.hh) -- one snapshot */
void cpo_gen_woven(int W, int H, double t, double *out);                         /* scalar */
void cpo_gen_merger(int W, int H, double t, double *out);                        /* scalar */
void cpo_gen_moving_extremum(int nd, const int32_t *dims, const double *x0, const double *dir, double t, double *out);
void cpo_gen_double_gyre(int W, int H, double time, double A, double omega, double eps, double *out); /* vector (2,W,H) */
void cpo_gen_abc(int W, int H, int D, double A, double B, double C, double *out); /* vector (3,W,H,D) */
void cpo_gen_tornado(int W, int H, int D, int time, double *out);                /* vector (3,W,H,D) */

#ifdef __cplusplus
}
#endif
#endif
