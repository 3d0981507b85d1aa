// See mesh_tables.h.  ref: include/ftk/mesh/simplicial_regular_mesh.hh:620-831 (semantics only).
#include "mesh_tables.h"

#include <algorithm>
#include <cstring>
#include <mutex>
#include <stdexcept>

namespace ftkb {
namespace {

// lexicographic rank of a vertex mask when dimension 0 is compared first
inline int lexkey(int mask, int nd) {
  int k = 0;
  for (int j = 0; j < nd; j++) k |= ((mask >> j) & 1) << (nd - 1 - j);
  return k;
}

void enumerate_chains(int nd, int k, std::vector<uint8_t> &cur, std::vector<std::vector<uint8_t>> &out) {
  if ((int)cur.size() == k + 1) {
    out.push_back(cur);
    return;
  }
  const int last = cur.back();
  for (int m = 1; m < (1 << nd); m++)
    if ((m & last) == last && m != last) {  // strict superset of the previous vertex
      cur.push_back((uint8_t)m);
      enumerate_chains(nd, k, cur, out);
      cur.pop_back();
    }
}

bool offset_less(const TypeOffset &a, const TypeOffset &b, int nd) {
  if (a.type != b.type) return a.type < b.type;
  for (int j = 0; j < nd; j++)
    if (a.off[j] != b.off[j]) return a.off[j] < b.off[j];
  return false;
}

int find_type(const MeshTables &m, int k, const std::vector<uint8_t> &chain) {
  const auto &types = m.unit[k];
  for (int t = 0; t < (int)types.size(); t++)
    if (types[t] == chain) return t;
  throw std::logic_error("ftkb: simplex type not found");
}

void build(MeshTables &m, int nd) {
  m.nd = nd;
  m.unit.resize(nd + 1);
  for (int k = 0; k <= nd; k++) {
    std::vector<uint8_t> cur{0};
    enumerate_chains(nd, k, cur, m.unit[k]);
    std::sort(m.unit[k].begin(), m.unit[k].end(), [nd](const std::vector<uint8_t> &a, const std::vector<uint8_t> &b) {
      for (size_t i = 0; i < a.size(); i++) {
        const int ka = lexkey(a[i], nd), kb = lexkey(b[i], nd);
        if (ka != kb) return ka < kb;
      }
      return false;
    });
  }
  // ordinal = no vertex is displaced in time
  m.ordinal_types.resize(nd + 1);
  m.interval_types.resize(nd + 1);
  m.is_ordinal.resize(nd + 1);
  const int tbit = 1 << (nd - 1);
  for (int k = 0; k <= nd; k++)
    for (int t = 0; t < m.ntypes(k); t++) {
      bool ord = true;
      for (uint8_t v : m.unit[k][t]) ord = ord && !(v & tbit);
      m.is_ordinal[k].push_back(ord);
      (ord ? m.ordinal_types[k] : m.interval_types[k]).push_back(t);
    }
  // facets: drop one vertex, translate so that the smallest vertex is the corner
  m.sides.resize(nd + 1);
  for (int k = 1; k <= nd; k++) {
    m.sides[k].resize(m.ntypes(k));
    for (int t = 0; t < m.ntypes(k); t++) {
      const auto &s = m.unit[k][t];
      for (int drop = 0; drop <= k; drop++) {
        std::vector<uint8_t> f;
        for (int i = 0; i <= k; i++)
          if (i != drop) f.push_back(s[i]);
        const uint8_t base = f[0];  // chain => f[0] is contained in every other vertex
        for (auto &v : f) v = (uint8_t)(v & ~base);
        TypeOffset e{};
        e.type = find_type(m, k - 1, f);
        for (int j = 0; j < nd; j++) e.off[j] = (base >> j) & 1;
        m.sides[k][t].push_back(e);
      }
      std::sort(m.sides[k][t].begin(), m.sides[k][t].end(),
                [nd](const TypeOffset &a, const TypeOffset &b) { return offset_less(a, b, nd); });
    }
  }
  m.sides[0].resize(m.ntypes(0));
  // cofaces: every (k+1)-type placed at every corner offset in {-1,0,1}^nd that contains the simplex
  m.side_of.resize(nd + 1);
  for (int k = 0; k < nd; k++) {
    m.side_of[k].resize(m.ntypes(k));
    int ncorners = 1;
    for (int j = 0; j < nd; j++) ncorners *= 3;
    for (int t = 0; t < m.ntypes(k); t++) {
      const auto &s = m.unit[k][t];
      for (int h = 0; h < m.ntypes(k + 1); h++) {
        const auto &hs = m.unit[k + 1][h];
        for (int ci = 0; ci < ncorners; ci++) {
          int off[4] = {0, 0, 0, 0};
          for (int j = 0, c = ci; j < nd; j++, c /= 3) off[j] = (c % 3) - 1;
          bool all = true;
          for (uint8_t v : s) {
            bool found = false;
            for (uint8_t w : hs) {
              bool eq = true;
              for (int j = 0; j < nd; j++) eq = eq && (((w >> j) & 1) + off[j] == ((v >> j) & 1));
              found = found || eq;
            }
            all = all && found;
          }
          if (all) {
            TypeOffset e{};
            e.type = h;
            std::memcpy(e.off, off, sizeof(off));
            m.side_of[k][t].push_back(e);
          }
        }
      }
      std::sort(m.side_of[k][t].begin(), m.side_of[k][t].end(),
                [nd](const TypeOffset &a, const TypeOffset &b) { return offset_less(a, b, nd); });
    }
  }
  m.side_of[nd].resize(m.ntypes(nd));
}

}  // namespace

const MeshTables &mesh_tables(int nd_mesh) {
  static MeshTables tables[2];
  static std::once_flag once[2];
  if (nd_mesh != 3 && nd_mesh != 4) throw std::invalid_argument("ftkb: mesh dimensionality must be 3 or 4");
  std::call_once(once[nd_mesh - 3], [nd_mesh]() { build(tables[nd_mesh - 3], nd_mesh); });
  return tables[nd_mesh - 3];
}

void fill_device_tables(int nd_mesh, DeviceMeshTables *out) {
  const MeshTables &m = mesh_tables(nd_mesh);
  const int n = nd_mesh - 1;
  std::memset(out, 0, sizeof(*out));
  out->nd = nd_mesh;
  out->ntypes = m.ntypes(n);
  if (out->ntypes > 60) throw std::logic_error("ftkb: too many simplex types");
  for (int t = 0; t < out->ntypes; t++) {
    uint8_t all = 0;
    for (int i = 0; i <= n; i++) {
      out->vmask[t][i] = m.unit[n][t][i];
      all |= m.unit[n][t][i];
    }
    out->tmask[t] = all;
    out->ordinal[t] = m.is_ordinal[n][t];
    // neighbours: the other facets of every coface (ref: critical_point_tracker_2d_regular.hh:189-197)
    int cnt = 0;
    for (const TypeOffset &cell : m.side_of[n][t])
      for (const TypeOffset &f : m.sides[nd_mesh][cell.type]) {
        int off[4] = {0, 0, 0, 0};
        bool self = f.type == t;
        for (int j = 0; j < nd_mesh; j++) {
          off[j] = cell.off[j] + f.off[j];
          self = self && off[j] == 0;
        }
        if (self) continue;
        bool dup = false;
        for (int q = 0; q < cnt; q++) {
          bool same = out->nb_type[t][q] == f.type;
          for (int j = 0; j < nd_mesh; j++) same = same && out->nb_off[t][q][j] == off[j];
          dup = dup || same;
        }
        if (dup) continue;
        if (cnt >= 8) throw std::logic_error("ftkb: more than 8 neighbour candidates");
        out->nb_type[t][cnt] = (int8_t)f.type;
        for (int j = 0; j < 4; j++) out->nb_off[t][cnt][j] = (int8_t)off[j];
        cnt++;
      }
    out->n_nb[t] = (uint8_t)cnt;
  }
}

}  // namespace ftkb
