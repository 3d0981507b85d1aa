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
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>

#include "kernels.h"

namespace ftkb {

OnlineTracer::OnlineTracer(int nd, const int32_t lb[3], const int32_t ub[3]) : nd_(nd) {
  for (int j = 0; j < 3; j++) {
    lb_[j] = j < nd ? lb[j] : 0;
    ub_[j] = j < nd ? ub[j] : 0;
  }
  ny_ = ub_[1] - lb_[1] + 1;
  nz_ = ub_[2] - lb_[2] + 1;
  DeviceMeshTables mt;
  fill_device_tables(nd + 1, &mt);
  ntypes_ = mt.ntypes;
  for (int type = 0; type < mt.ntypes; type++) {
    int cnt = 0;
    Candidate self{};
    self.type = (int8_t)type;
    cand_[type][cnt++] = self;
    for (int q = 0; q < mt.n_nb[type]; q++) {
      Candidate c{};
      c.off[0] = mt.nb_off[type][q][0];
      c.off[1] = mt.nb_off[type][q][1];
      c.off[2] = nd == 3 ? mt.nb_off[type][q][2] : 0;
      c.off[3] = mt.nb_off[type][q][nd];
      c.type = mt.nb_type[type][q];
      cand_[type][cnt++] = c;
    }
    std::sort(cand_[type], cand_[type] + cnt, [](const Candidate &a, const Candidate &b) {
      for (int j = 0; j < 4; j++) if (a.off[j] != b.off[j]) return a.off[j] < b.off[j];
      return a.type < b.type;
    });
    for (int q = 0; q < cnt; q++) {
      Candidate &c = cand_[type][q];
      c.cell_delta = ((int64_t)c.off[0] * ny_ + c.off[1]) * nz_ + c.off[2];
      // key = ((cell << TIME) | t) << TYPE | type is linear in (cell, t, type) as long as no field leaves its range
      c.key_delta = c.cell_delta * ((int64_t)1 << (KEY_TIME_BITS + KEY_TYPE_BITS)) + (int64_t)c.off[3] * ((int64_t)1 << KEY_TYPE_BITS) + ((int64_t)c.type - type);
    }
    ncand_[type] = cnt;
  }
}

bool OnlineTracer::key_at(int x, int y, int z, int t, int type, uint64_t &key) const {
  if (x < lb_[0] || x > ub_[0] || y < lb_[1] || y > ub_[1] || t < 0 || t >= (1 << KEY_TIME_BITS)) return false;
  if (nd_ == 3 && (z < lb_[2] || z > ub_[2])) return false;
  if (type < 0 || type >= ntypes_) return false;
  uint64_t k = (uint64_t)(x - lb_[0]);
  k = k * (uint64_t)ny_ + (uint64_t)(y - lb_[1]);
  k = k * (uint64_t)nz_ + (uint64_t)(nd_ == 3 ? z - lb_[2] : 0);
  k = (k << KEY_TIME_BITS) | (uint64_t)t;
  key = (k << KEY_TYPE_BITS) | (uint64_t)type;
  return true;
}

// neighbors(f) = every side of every cell f is a side of (critical_point_tracker_2d_regular.hh:292-300), f included;
// computed on the element's key alone
int OnlineTracer::neighbor_keys(const uint64_t self, uint64_t out[9]) const {
  const int type = (int)(self & ((1u << KEY_TYPE_BITS) - 1));
  if (type >= ntypes_) return 0;
  const int t0 = (int)((self >> KEY_TYPE_BITS) & ((1u << KEY_TIME_BITS) - 1));
  const int64_t cell = (int64_t)(self >> (KEY_TYPE_BITS + KEY_TIME_BITS));
  const int z0 = (int)(cell % nz_) + lb_[2], y0 = (int)((cell / nz_) % ny_) + lb_[1], x0 = (int)(cell / (nz_ * ny_)) + lb_[0];
  if (x0 > lb_[0] && x0 < ub_[0] && y0 > lb_[1] && y0 < ub_[1] && (nd_ == 2 || (z0 > lb_[2] && z0 < ub_[2])) && t0 > 0 && t0 < (1 << KEY_TIME_BITS) - 1) {
    // away from the domain's faces every candidate exists (offsets are -1, 0, 1): its key is the element's plus a constant
    const int n = ncand_[type];
    for (int q = 0; q < n; q++) out[q] = self + (uint64_t)cand_[type][q].key_delta;
    return n;
  }
  int cnt = 0;
  for (int q = 0; q < ncand_[type]; q++) {
    const Candidate &c = cand_[type][q];
    const int x = x0 + c.off[0], y = y0 + c.off[1], z = z0 + c.off[2], t = t0 + c.off[3];
    if (x < lb_[0] || x > ub_[0] || y < lb_[1] || y > ub_[1] || z < lb_[2] || z > ub_[2] || t < 0 || t >= (1 << KEY_TIME_BITS)) continue;
    out[cnt++] = ((((uint64_t)(cell + c.cell_delta) << KEY_TIME_BITS) | (uint64_t)t) << KEY_TYPE_BITS) | (uint64_t)c.type;
  }
  return cnt;
}

const ftkb_point &OnlineTracer::point(uint32_t g) const {
  const size_t b = (size_t)(std::upper_bound(batch_base_.begin(), batch_base_.end(), g) - batch_base_.begin()) - 1;
  return batches_[b][g - batch_base_[b]];
}

uint64_t OnlineTracer::npoints() const {
  uint64_t n = 0;
  for (const OnlineCurve &c : curves_) n += c.idx.size();
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
// Components come back ordered by root, members ascending (union_find.hh:82-92), as CSR: component g holds
// members[start[g] .. start[g+1]).  `local` is scratch of the universe's size, all -1 on entry and on exit.
struct Components {
  std::vector<uint32_t> start, members;
  size_t size() const { return start.size() - 1; }
};

template <class NB>
Components components_of(const std::vector<uint32_t> &nodes, std::vector<int32_t> &local, NB nb) {
  const size_t m = nodes.size();
  for (size_t a = 0; a < m; a++) local[nodes[a]] = (int32_t)a;
  RefUnionFind uf(m);
  uint32_t tmp[9];
  for (size_t a = 0; a < m; a++) {
    const int cnt = nb(nodes[a], tmp);
    for (int q = 0; q < cnt; q++)
      if (local[tmp[q]] >= 0) uf.unite((uint32_t)a, (uint32_t)local[tmp[q]]);
  }
  // get_sets(): find() of every element in ascending order (it compresses paths as it goes), sets ordered by root
  std::vector<uint32_t> root(m), rank(m, 0);
  for (size_t a = 0; a < m; a++) { root[a] = uf.find((uint32_t)a); rank[root[a]] = 1; }
  Components out;
  uint32_t ngroups = 0;
  for (size_t a = 0; a < m; a++) { const uint32_t is_root = rank[a]; rank[a] = ngroups; ngroups += is_root; }
  out.start.assign(ngroups + 1, 0);
  for (size_t a = 0; a < m; a++) out.start[rank[root[a]] + 1]++;
  for (uint32_t g = 0; g < ngroups; g++) out.start[g + 1] += out.start[g];
  out.members.resize(m);
  std::vector<uint32_t> fill(out.start.begin(), out.start.end() - 1);
  for (size_t a = 0; a < m; a++) out.members[fill[rank[root[a]]]++] = nodes[a];
  for (size_t a = 0; a < m; a++) local[nodes[a]] = -1;
  return out;
}

}  // namespace

void OnlineTracer::grow(const ftkb_point *pts_in, uint64_t n_in) {
  // discrete_critical_points: std::map keyed by element (a later insert of the same element overwrites).  Most of a
  // step's elements are claimed by existing trajectories through key lookups alone, so the batch is hashed, not
  // sorted; only what is left for step 2 is put into element order.
  size_t cap = 16;
  while (cap < 2 * n_in + 2) cap <<= 1;
  int shift = 64;
  for (size_t c = cap; c > 1; c >>= 1) shift--;
  // open addressing, hashed on the element's (corner, t) without the type: the neighbours of an element sit in a
  // handful of adjacent cells, so their probes share cache lines; probes of one neighbour list are prefetched together
  struct Entry { uint64_t key; uint32_t val, pad; };
  std::vector<Entry> table(cap, Entry{~0ull, 0, 0});   // no element key is all ones (the type field is below 60)
  std::vector<uint64_t> keys;
  keys.reserve(n_in);
  std::vector<ftkb_point> batch;
  batch.reserve(n_in);
  auto home_of = [&](uint64_t key) { return (size_t)(((key >> KEY_TYPE_BITS) * 0x9E3779B97F4A7C15ull) >> shift); };
  auto slot_from = [&](size_t h, uint64_t key) {
    while (table[h].key != ~0ull && table[h].key != key) h = (h + 1) & (cap - 1);
    return h;
  };
  for (uint64_t i = 0; i < n_in; i++) {
    uint64_t key;
    if (!key_of(pts_in[i], key)) continue;
    const size_t h = slot_from(home_of(key), key);
    if (table[h].key == key) { batch[table[h].val] = pts_in[i]; continue; }
    table[h].key = key;
    table[h].val = (uint32_t)keys.size();
    keys.push_back(key);
    batch.push_back(pts_in[i]);
  }
  // the punctured neighbours of an element, ascending (the candidate order), as batch indices; looked up when needed
  struct Source {
    const OnlineTracer *self; const uint64_t *keys; const std::vector<Entry> *table; size_t cap; int shift;
    size_t home_of(uint64_t key) const { return (size_t)(((key >> KEY_TYPE_BITS) * 0x9E3779B97F4A7C15ull) >> shift); }
    size_t slot_from(size_t h, uint64_t key) const {
      const Entry *t = table->data();
      while (t[h].key != ~0ull && t[h].key != key) h = (h + 1) & (cap - 1);
      return h;
    }
    int64_t find(uint64_t key) const {
      const size_t h = slot_from(home_of(key), key);
      return (*table)[h].key == key ? (int64_t)(*table)[h].val : -1;
    }
    int list(uint32_t i, uint32_t out[9]) const {
      uint64_t nk[9];
      size_t home[9];
      const int m = self->neighbor_keys(keys[i], nk);
      for (int q = 0; q < m; q++) { home[q] = home_of(nk[q]); __builtin_prefetch(table->data() + home[q]); }
      int cnt = 0;
      for (int q = 0; q < m; q++) {
        const size_t h = slot_from(home[q], nk[q]);
        if ((*table)[h].key == nk[q]) out[cnt++] = (*table)[h].val;
      }
      return cnt;
    }
    // first live entry of the list: probing stops at the first hit
    int64_t first_alive(uint32_t i, const uint8_t *alive) const {
      uint64_t nk[9];
      size_t home[9];
      const int m = self->neighbor_keys(keys[i], nk);
      for (int q = 0; q < m; q++) { home[q] = home_of(nk[q]); __builtin_prefetch(table->data() + home[q]); }
      for (int q = 0; q < m; q++) {
        const size_t h = slot_from(home[q], nk[q]);
        if ((*table)[h].key == nk[q] && alive[(*table)[h].val]) return (int64_t)(*table)[h].val;
      }
      return -1;
    }
  };
  store_batch(keys.data(), (uint32_t)keys.size(), std::move(batch));
  walk((uint32_t)keys.size(), keys.data(), false, Source{this, keys.data(), &table, cap, shift});
}

// every element handed to a grow step: its key (the walk needs nothing else of an earlier step's element) and, for callers that
// read trajectories back as points, the batch's records -- one array per batch, never copied again
void OnlineTracer::store_batch(const uint64_t *keys, uint32_t n, std::vector<ftkb_point> &&pts) {
  if (keys_.size() + n >= 0xffffffffull) throw std::length_error("more than 2^32 punctured simplices");
  batch_base_.push_back((uint32_t)keys_.size());
  keys_.insert(keys_.end(), keys, keys + n);
  batches_.push_back(std::move(pts));
}

void OnlineTracer::grow_sorted(const ftkb_point *pts, const uint64_t *keys, uint32_t n, const uint32_t *nb, const uint8_t *cnt) {
  if (!pts && n) have_points_ = false;
  store_batch(keys, n, pts ? std::vector<ftkb_point>(pts, pts + n) : std::vector<ftkb_point>());
  struct Source {
    const uint64_t *keys; uint32_t n; const uint32_t *nb; const uint8_t *cnt;
    int64_t find(uint64_t key) const {
      const uint64_t *it = std::lower_bound(keys, keys + n, key);
      return it != keys + n && *it == key ? (int64_t)(it - keys) : -1;
    }
    int list(uint32_t i, uint32_t out[9]) const { std::memcpy(out, nb + 9 * (size_t)i, 4 * cnt[i]); return cnt[i]; }
    int64_t first_alive(uint32_t i, const uint8_t *alive) const {
      const uint32_t *l = nb + 9 * (size_t)i;
      for (int q = 0; q < cnt[i]; q++) if (alive[l[q]]) return (int64_t)l[q];
      return -1;
    }
  };
  walk(n, keys, true, Source{keys, n, nb, cnt});
}

// The batch is the last n stored elements.
template <class Source>
void OnlineTracer::walk(uint32_t n, const uint64_t *keys, bool sorted, Source src) {
  const uint32_t base = (uint32_t)(keys_.size() - n);
  std::vector<uint8_t> alive(n, 1);
  // smallest unclaimed punctured neighbour of an element (critical_point_tracker.hh:555-572): the first live entry of its
  // ascending neighbour list.  An element of this batch has its list ready; a trajectory's end from an earlier step looks
  // its candidates up by key (only the first claim of an end does).
  auto claim_next = [&](uint32_t g) -> int64_t {
    if (g >= base) {
      const int64_t j = src.first_alive(g - base, alive.data());
      if (j >= 0) alive[j] = 0;
      return j;
    }
    if (n == 0) return -1;
    uint64_t nk[9];
    const int m = neighbor_keys(keys_[g], nk);
    for (int q = 0; q < m; q++) {
      const int64_t j = src.find(nk[q]);
      if (j >= 0 && alive[j]) { alive[j] = 0; return j; }
    }
    return -1;
  };

  // 1. continue existing trajectories, in id order (critical_point_tracker.hh:540-609)
  for (OnlineCurve &c : curves_) {
    if (c.complete || c.idx.empty()) continue;
    bool continued = false;
    for (uint32_t cur = c.idx.back();;) {
      const int64_t j = claim_next(cur);
      if (j < 0) break;
      cur = base + (uint32_t)j;
      c.idx.push_back(cur);
      continued = true;
    }
    for (uint32_t cur = c.idx.front();;) {
      const int64_t j = claim_next(cur);
      if (j < 0) break;
      cur = base + (uint32_t)j;
      c.idx.push_front(cur);
      continued = true;
    }
    if (!continued) c.complete = true;
  }

  // 2. new trajectories from what is left (critical_point_tracker.hh:611-640), in std::set<element> order
  std::vector<uint32_t> rest;
  for (uint32_t i = 0; i < n; i++) if (alive[i]) rest.push_back(i);
  if (rest.empty()) return;
  if (!sorted) std::sort(rest.begin(), rest.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
  // neighbour lists restricted to what is left
  std::vector<uint32_t> lnb(9 * (size_t)n);
  std::vector<uint8_t> nnb(n, 0);
  for (uint32_t i : rest) {
    uint32_t l[9];
    const int m = src.list(i, l);
    for (int q = 0; q < m; q++)
      if (alive[l[q]]) lnb[9 * (size_t)i + nnb[i]++] = l[q];
  }
  auto nb_all = [&](uint32_t i, uint32_t *out) { std::memcpy(out, &lnb[9 * (size_t)i], 4 * nnb[i]); return (int)nnb[i]; };
  std::vector<int32_t> scratch(n, -1);
  const Components components = components_of(rest, scratch, nb_all);

  // cc2curves.hh:19-31: a node with more than two punctured neighbours other than itself is special.  The reference
  // splits every component into linear graphs on its own; unions never cross components and every component's
  // ordinary nodes are visited in the same relative order, so one union-find over all ordinary nodes (cc2curves.hh:33-43)
  // gives the same sets, and ordering them by (component, root) gives the reference's order of new trajectories.
  std::vector<uint8_t> special(n, 0), visited(n, 0);
  std::vector<uint32_t> comp_of(n, 0), ordinary;
  ordinary.reserve(rest.size());
  for (uint32_t g = 0; g < components.size(); g++)
    for (uint32_t k = components.start[g]; k < components.start[g + 1]; k++) comp_of[components.members[k]] = g;
  for (uint32_t i : rest) {
    int d = 0;
    for (int q = 0; q < nnb[i]; q++) d += lnb[9 * (size_t)i + q] != i;
    special[i] = d > 2;
    if (!special[i]) ordinary.push_back(i);
  }
  auto nb_ord = [&](uint32_t i, uint32_t *out) {
    int m = 0;
    for (int q = 0; q < nnb[i]; q++) { const uint32_t j = lnb[9 * (size_t)i + q]; if (!special[j]) out[m++] = j; }
    return m;
  };
  const Components linear = components_of(ordinary, scratch, nb_ord);
  std::vector<int32_t> member(n, -1);   // index of the linear graph a node belongs to
  std::vector<uint32_t> graph_order(linear.size());
  for (uint32_t g = 0; g < linear.size(); g++) {
    graph_order[g] = g;
    for (uint32_t k = linear.start[g]; k < linear.start[g + 1]; k++) member[linear.members[k]] = (int32_t)g;
  }
  std::stable_sort(graph_order.begin(), graph_order.end(), [&](uint32_t a, uint32_t b) {
    return comp_of[linear.members[linear.start[a]]] < comp_of[linear.members[linear.start[b]]];
  });
  // cc2curves.hh:46-108
  std::vector<uint32_t> fwd, bwd;
  for (const uint32_t g : graph_order) {
    const uint32_t seed = linear.members[linear.start[g]];
    fwd.clear();
    bwd.clear();
    visited[seed] = 1;
    uint32_t sn[9]; int nsn = 0;
    for (int q = 0; q < nnb[seed]; q++) { const uint32_t j = lnb[9 * (size_t)seed + q]; if (j != seed && !special[j]) sn[nsn++] = j; }
    for (int dir = 0; dir < 2 && nsn > 0; dir++) {
      uint32_t cur = dir == 0 ? sn[0] : sn[nsn - 1];
      while (true) {
        if (!visited[cur]) {
          (dir == 0 ? fwd : bwd).push_back(cur);
          visited[cur] = 1;
        }
        bool found = false;
        for (int q = 0; q < nnb[cur]; q++) {
          const uint32_t j = lnb[9 * (size_t)cur + q];
          if (j != cur && !special[j] && member[j] == (int32_t)g && !visited[j]) { found = true; cur = j; break; }
        }
        if (!found) break;
      }
      if (nsn == 1) break;
    }
    OnlineCurve c;
    const uint32_t front = bwd.empty() ? seed : bwd.back(), back = fwd.empty() ? seed : fwd.back();
    if (fwd.size() + bwd.size() > 0)   // is_loop, cc2curves.hh:113-122
      for (int q = 0; q < nnb[front]; q++) c.loop = c.loop || lnb[9 * (size_t)front + q] == back;
    for (size_t k = bwd.size(); k > 0; k--) c.idx.push_back(base + bwd[k - 1]);
    c.idx.push_back(base + seed);
    for (uint32_t i : fwd) c.idx.push_back(base + i);
    curves_.push_back(std::move(c));
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

// the same step through grow_sorted: the preparation the tracker does on the device (sort by element key, one entry per
// element, neighbour lists by binary search, the element itself included) restated on the host, so that the index-list walk is
// tested without a device
extern "C" int ftkb_online_grow_prepared(ftkb_online *o, const ftkb_point *pts, uint64_t n) {
  if (!o || (!pts && n) || n >= 0xffffffffull) return FTKB_ERR_INVALID;
  std::vector<std::pair<uint64_t, uint32_t>> order;
  order.reserve(n);
  for (uint64_t i = 0; i < n; i++) {
    uint64_t key;
    if (o->tracer.key_of(pts[i], key)) order.emplace_back(key, (uint32_t)i);
  }
  std::stable_sort(order.begin(), order.end(), [](const std::pair<uint64_t, uint32_t> &a, const std::pair<uint64_t, uint32_t> &b) { return a.first < b.first; });
  std::vector<ftkb_point> sorted;
  std::vector<uint64_t> keys;
  for (size_t k = 0; k < order.size(); k++)
    if (k == 0 || order[k].first != order[k - 1].first) { keys.push_back(order[k].first); sorted.push_back(pts[order[k].second]); }
  const uint32_t m = (uint32_t)keys.size();
  std::vector<uint32_t> nb(9 * (size_t)m, 0xffffffffu);
  std::vector<uint8_t> cnt(m, 0);
  for (uint32_t i = 0; i < m; i++) {
    uint64_t nk[9];
    const int c = o->tracer.neighbor_keys(keys[i], nk);     // ascending
    for (int q = 0; q < c; q++) {
      const auto it = std::lower_bound(keys.begin(), keys.end(), nk[q]);
      if (it != keys.end() && *it == nk[q]) nb[9 * (size_t)i + cnt[i]++] = (uint32_t)(it - keys.begin());
    }
  }
  static const bool timing = std::getenv("FTKB_DEBUG_TIMING") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  o->tracer.grow_sorted(sorted.data(), keys.data(), m, nb.data(), cnt.data());
  if (timing) {
    static double total = 0;
    total += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr, "[ftkb timing] grow_sorted walk alone: %.3f ms so far\n", total);
  }
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
  if (pts && !o->tracer.has_points()) return FTKB_ERR_INVALID;
  for (const ftkb::OnlineCurve &c : o->tracer.curves()) {
    if (pts) for (const uint32_t g : c.idx) pts[pos++] = o->tracer.point(g); else pos += c.idx.size();
    if (loop) loop[k] = c.loop;
    if (complete) complete[k] = c.complete;
    offsets[++k] = pos;
  }
  return FTKB_OK;
}
