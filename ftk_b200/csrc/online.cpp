// Streaming trajectories: host code, no device work.
//
// The reference grows its trajectories after every interval sweep with a greedy, order-dependent walk
// (critical_point_tracker.hh:522-641, called from critical_point_tracker_2d_regular.hh:288-329 and
// ..._3d_regular.hh:173-200).  The batch of one step is a few punctured simplices per feature, and every decision
// depends on the one before it, so the walk stays on the host; the device produced and compacted the batch.
// What is restated here, in the reference's iteration orders (std::set / std::map of elements = ascending element key):
//   1. every incomplete trajectory is extended forwards from its back and backwards from its front by the smallest
//      punctured neighbour still unclaimed, until none is left; a trajectory that took nothing becomes complete;
//   2. the unclaimed rest is split into connected components (algorithms/cca.hh:91-116 over basic/union_find.hh),
//      each component into linear graphs (geometry/cc2curves.hh:10-122), each linear graph becomes a new trajectory.
// The component ORDER decides trajectory ids and, in later steps, who claims a shared neighbour first; it follows
// from union_find's weighted union including its quirks (no same-set check: uniting an element with itself or with
// a member of its own set doubles the recorded size), so those are kept as they are.
#include "online.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <new>

#include "kernels.h"

namespace ftkb {

OnlineTracer::OnlineTracer(int nd, const int32_t lb[3], const int32_t ub[3]) : nd_(nd) {
  for (int j = 0; j < 3; j++) {
    lb_[j] = j < nd ? lb[j] : 0;
    ub_[j] = j < nd ? ub[j] : 0;
  }
  ny_ = ub_[1] - lb_[1] + 1;
  nz_ = ub_[2] - lb_[2] + 1;
  fill_device_tables(nd + 1, &mt_);
}

bool OnlineTracer::key_at(int x, int y, int z, int t, int type, uint64_t &key) const {
  if (x < lb_[0] || x > ub_[0] || y < lb_[1] || y > ub_[1] || t < 0 || t >= (1 << KEY_TIME_BITS)) return false;
  if (nd_ == 3 && (z < lb_[2] || z > ub_[2])) return false;
  uint64_t k = (uint64_t)(x - lb_[0]);
  k = k * (uint64_t)ny_ + (uint64_t)(y - lb_[1]);
  k = k * (uint64_t)nz_ + (uint64_t)(nd_ == 3 ? z - lb_[2] : 0);
  k = (k << KEY_TIME_BITS) | (uint64_t)t;
  key = (k << KEY_TYPE_BITS) | (uint64_t)type;
  return true;
}

// neighbors(f) = every side of every cell f is a side of (critical_point_tracker_2d_regular.hh:292-300), f included
int OnlineTracer::neighbor_keys(const ftkb_point &p, uint64_t out[9]) const {
  int cnt = 0;
  const int type = p.simplex_type;
  uint64_t key;
  if (key_of(p, key)) out[cnt++] = key;
  for (int q = 0; q < mt_.n_nb[type]; q++) {
    const int x = p.corner[0] + mt_.nb_off[type][q][0], y = p.corner[1] + mt_.nb_off[type][q][1];
    const int z = nd_ == 3 ? p.corner[2] + mt_.nb_off[type][q][2] : 0;
    const int t = p.corner[3] + mt_.nb_off[type][q][nd_];
    if (key_at(x, y, z, t, mt_.nb_type[type][q], key)) out[cnt++] = key;
  }
  std::sort(out, out + cnt);
  return cnt;
}

uint64_t OnlineTracer::npoints() const {
  uint64_t n = 0;
  for (const OnlineCurve &c : curves_) n += c.pts.size();
  return n;
}

namespace {

// basic/union_find.hh:15-107 over dense indices (index order = element order)
struct RefUnionFind {
  std::vector<uint32_t> parent;
  std::vector<uint64_t> sz;
  explicit RefUnionFind(size_t n) : parent(n), sz(n, 1) { for (size_t i = 0; i < n; i++) parent[i] = (uint32_t)i; }
  // union_find.hh:57-68: `parent_i` is a reference to the ORIGINAL element's slot, so the path update writes there
  uint32_t find(uint32_t i) {
    const uint32_t orig = i;
    while (i != parent[orig]) {
      parent[i] = parent[parent[orig]];
      i = parent[i];
      parent[orig] = parent[i];
    }
    return i;
  }
  // union_find.hh:32-45: weighted, no same-set check
  void unite(uint32_t i, uint32_t j) {
    i = find(i);
    j = find(j);
    if (sz[i] < sz[j]) { parent[i] = j; sz[j] += sz[i]; }
    else { parent[j] = i; sz[i] += sz[j]; }
  }
};

// algorithms/cca.hh:91-116: `nodes` ascending; nb(node) = its neighbours among `nodes` (ascending, itself included).
// Components come back ordered by root, members ascending (union_find.hh:82-92).
template <class NB>
std::vector<std::vector<uint32_t>> components_of(const std::vector<uint32_t> &nodes, uint32_t universe, NB nb) {
  std::vector<int32_t> local(universe, -1);
  for (size_t a = 0; a < nodes.size(); a++) local[nodes[a]] = (int32_t)a;
  RefUnionFind uf(nodes.size());
  uint32_t tmp[9];
  for (size_t a = 0; a < nodes.size(); a++) {
    const int cnt = nb(nodes[a], tmp);
    for (int q = 0; q < cnt; q++)
      if (local[tmp[q]] >= 0) uf.unite((uint32_t)a, (uint32_t)local[tmp[q]]);
  }
  std::map<uint32_t, std::vector<uint32_t>> root2set;
  for (size_t a = 0; a < nodes.size(); a++) root2set[uf.find((uint32_t)a)].push_back(nodes[a]);
  std::vector<std::vector<uint32_t>> out;
  out.reserve(root2set.size());
  for (auto &kv : root2set) out.push_back(std::move(kv.second));
  return out;
}

}  // namespace

void OnlineTracer::grow(const ftkb_point *pts_in, uint64_t n_in) {
  // discrete_critical_points: std::map keyed by element (a later insert of the same element overwrites)
  std::vector<std::pair<uint64_t, uint32_t>> order;
  order.reserve(n_in);
  for (uint64_t i = 0; i < n_in; i++) {
    uint64_t key;
    if (key_of(pts_in[i], key)) order.emplace_back(key, (uint32_t)i);
  }
  std::stable_sort(order.begin(), order.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
  std::vector<uint64_t> keys;
  std::vector<ftkb_point> pts;
  for (size_t a = 0; a < order.size(); a++) {
    if (!keys.empty() && keys.back() == order[a].first) { pts.back() = pts_in[order[a].second]; continue; }
    keys.push_back(order[a].first);
    pts.push_back(pts_in[order[a].second]);
  }
  const uint32_t n = (uint32_t)keys.size();
  std::vector<uint8_t> alive(n, 1);
  auto lookup = [&](uint64_t key) -> int64_t {
    const auto it = std::lower_bound(keys.begin(), keys.end(), key);
    return it != keys.end() && *it == key ? (int64_t)(it - keys.begin()) : -1;
  };
  // smallest unclaimed punctured neighbour of an element (critical_point_tracker.hh:555-572)
  auto claim_next = [&](const ftkb_point &cur) -> int64_t {
    uint64_t nk[9];
    const int cnt = neighbor_keys(cur, nk);
    for (int q = 0; q < cnt; q++) {
      const int64_t j = lookup(nk[q]);
      if (j >= 0 && alive[j]) { alive[j] = 0; return j; }
    }
    return -1;
  };

  // 1. continue existing trajectories, in id order (critical_point_tracker.hh:540-609)
  for (OnlineCurve &c : curves_) {
    if (c.complete || c.pts.empty()) continue;
    bool continued = false;
    for (ftkb_point cur = c.pts.back();;) {
      const int64_t j = claim_next(cur);
      if (j < 0) break;
      c.pts.push_back(pts[j]);
      cur = pts[j];
      continued = true;
    }
    for (ftkb_point cur = c.pts.front();;) {
      const int64_t j = claim_next(cur);
      if (j < 0) break;
      c.pts.push_front(pts[j]);
      cur = pts[j];
      continued = true;
    }
    if (!continued) c.complete = true;
  }

  // 2. new trajectories from what is left (critical_point_tracker.hh:611-640)
  std::vector<uint32_t> rest;
  for (uint32_t i = 0; i < n; i++) if (alive[i]) rest.push_back(i);
  if (rest.empty()) return;
  std::vector<uint32_t> nb(9 * (size_t)n, 0);
  std::vector<uint8_t> nnb(n, 0);
  for (uint32_t i : rest) {
    uint64_t nk[9];
    const int cnt = neighbor_keys(pts[i], nk);
    for (int q = 0; q < cnt; q++) {
      const int64_t j = lookup(nk[q]);
      if (j >= 0 && alive[j]) nb[9 * (size_t)i + nnb[i]++] = (uint32_t)j;
    }
  }
  auto nb_all = [&](uint32_t i, uint32_t *out) { std::memcpy(out, &nb[9 * (size_t)i], 4 * nnb[i]); return (int)nnb[i]; };
  const auto components = components_of(rest, n, nb_all);

  std::vector<uint8_t> special(n, 0), visited(n, 0);
  std::vector<int32_t> member(n, -1);   // index of the linear graph a node belongs to
  for (const auto &component : components) {
    // cc2curves.hh:19-31: more than two punctured neighbours other than itself
    std::vector<uint32_t> ordinary;
    for (uint32_t i : component) {
      int d = 0;
      for (int q = 0; q < nnb[i]; q++) d += nb[9 * (size_t)i + q] != i;
      special[i] = d > 2;
      if (!special[i]) ordinary.push_back(i);
    }
    // cc2curves.hh:33-43
    auto nb_ord = [&](uint32_t i, uint32_t *out) {
      int cnt = 0;
      for (int q = 0; q < nnb[i]; q++) { const uint32_t j = nb[9 * (size_t)i + q]; if (!special[j]) out[cnt++] = j; }
      return cnt;
    };
    const auto linear = components_of(ordinary, n, nb_ord);
    for (size_t g = 0; g < linear.size(); g++) for (uint32_t i : linear[g]) member[i] = (int32_t)g;
    // cc2curves.hh:46-108
    for (size_t g = 0; g < linear.size(); g++) {
      const uint32_t seed = linear[g].front();
      std::deque<uint32_t> trace;
      visited[seed] = 1;
      trace.push_back(seed);
      uint32_t sn[9]; int nsn = 0;
      for (int q = 0; q < nnb[seed]; q++) { const uint32_t j = nb[9 * (size_t)seed + q]; if (j != seed && !special[j]) sn[nsn++] = j; }
      for (int dir = 0; dir < 2 && nsn > 0; dir++) {
        uint32_t cur = dir == 0 ? sn[0] : sn[nsn - 1];
        while (true) {
          if (!visited[cur]) {
            if (dir == 0) trace.push_back(cur); else trace.push_front(cur);
            visited[cur] = 1;
          }
          bool found = false;
          for (int q = 0; q < nnb[cur]; q++) {
            const uint32_t j = nb[9 * (size_t)cur + q];
            if (j != cur && !special[j] && member[j] == (int32_t)g && !visited[j]) { found = true; cur = j; break; }
          }
          if (!found) break;
        }
        if (nsn == 1) break;
      }
      OnlineCurve c;
      if (trace.size() > 1) {   // is_loop, cc2curves.hh:113-122
        const uint32_t front = trace.front(), back = trace.back();
        for (int q = 0; q < nnb[front]; q++) c.loop = c.loop || nb[9 * (size_t)front + q] == back;
      }
      for (uint32_t i : trace) c.pts.push_back(pts[i]);
      curves_.push_back(std::move(c));
    }
    for (uint32_t i : component) member[i] = -1;
  }
}

}  // namespace ftkb

// ---- C ABI: the grow step by itself (host only) -----------------------------------------------------
struct ftkb_online {
  ftkb::OnlineTracer tracer;
  ftkb_online(int nd, const int32_t lb[3], const int32_t ub[3]) : tracer(nd, lb, ub) {}
};

extern "C" int ftkb_online_create(int nd, const int32_t *lb, const int32_t *ub, ftkb_online **out) {
  if (!out || !lb || !ub || (nd != 2 && nd != 3)) return FTKB_ERR_INVALID;
  for (int j = 0; j < nd; j++) if (ub[j] < lb[j]) return FTKB_ERR_INVALID;
  int32_t l[3] = {0, 0, 0}, u[3] = {0, 0, 0};
  for (int j = 0; j < nd; j++) { l[j] = lb[j]; u[j] = ub[j]; }
  *out = new (std::nothrow) ftkb_online(nd, l, u);
  return *out ? FTKB_OK : FTKB_ERR_NOMEM;
}

extern "C" void ftkb_online_destroy(ftkb_online *o) { delete o; }

extern "C" int ftkb_online_grow(ftkb_online *o, const ftkb_point *pts, uint64_t n) {
  if (!o || (!pts && n) || n >= 0xffffffffull) return FTKB_ERR_INVALID;
  o->tracer.grow(pts, n);
  return FTKB_OK;
}

extern "C" int ftkb_online_size(const ftkb_online *o, uint64_t *ntraj, uint64_t *npoints) {
  if (!o) return FTKB_ERR_INVALID;
  if (ntraj) *ntraj = o->tracer.curves().size();
  if (npoints) *npoints = o->tracer.npoints();
  return FTKB_OK;
}

extern "C" int ftkb_online_get(const ftkb_online *o, uint64_t *offsets, ftkb_point *pts, uint8_t *loop, uint8_t *complete) {
  if (!o || !offsets) return FTKB_ERR_INVALID;
  uint64_t pos = 0, k = 0;
  offsets[0] = 0;
  for (const ftkb::OnlineCurve &c : o->tracer.curves()) {
    if (pts) for (const ftkb_point &p : c.pts) pts[pos++] = p; else pos += c.pts.size();
    if (loop) loop[k] = c.loop;
    if (complete) complete[k] = c.complete;
    offsets[++k] = pos;
  }
  return FTKB_OK;
}
