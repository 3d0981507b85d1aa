// Trajectory post-processing on the host; see curves.h for the reference lines each part follows.
#include "curves.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>

namespace {

typedef ftkb_curveset::Curve Curve;

void relabel(Curve &c, int id) {
  c.id = id;
  for (auto &q : c.pts) q.id = id;
}

double vmag(const ftkb_curve_point &q) { return std::sqrt(q.v[0] * q.v[0] + q.v[1] * q.v[1] + q.v[2] * q.v[2]); }

// feature_curve.hh:145-186
void update_statistics(Curve &c) {
  if (c.pts.empty()) return;
  const double lo = std::numeric_limits<double>::lowest(), hi = std::numeric_limits<double>::max();
  c.smax = lo; c.smin = hi; c.tmax = lo; c.tmin = hi; c.vmmax = lo; c.vmmin = hi;
  for (int k = 0; k < 3; k++) { c.bbmax[k] = lo; c.bbmin[k] = hi; }
  for (const auto &q : c.pts) {
    c.smax = std::max(c.smax, q.p.scalar); c.smin = std::min(c.smin, q.p.scalar);
    for (int k = 0; k < 3; k++) { c.bbmax[k] = std::max(c.bbmax[k], q.p.x[k]); c.bbmin[k] = std::min(c.bbmin[k], q.p.x[k]); }
    c.tmax = std::max(c.tmax, q.p.t); c.tmin = std::min(c.tmin, q.p.t);
    c.vmmax = std::max(c.vmmax, vmag(q)); c.vmmin = std::min(c.vmmin, vmag(q));
  }
  c.persistence = c.smax - c.smin;
  c.consistent_type = c.pts[0].p.cp_type;
  for (const auto &q : c.pts)
    if (q.p.cp_type != c.consistent_type) { c.consistent_type = 0; break; }
}

std::vector<int> to_ordinals(const Curve &c) {
  std::vector<int> r;
  for (int i = 0; i < (int)c.pts.size(); i++)
    if (c.pts[i].p.ordinal) r.push_back(i);
  return r;
}

// feature_curve.hh:303-325 (half window 2)
void smooth_ordinal_types(Curve &c, const int hw = 2) {
  const auto ord = to_ordinals(c);
  if ((int)ord.size() < hw * 2 + 1) return;
  std::map<int, uint32_t> pending;
  for (int i = hw; i < (int)ord.size() - hw; i++) {
    uint32_t local = c.pts[ord[i - hw]].p.cp_type;
    for (int j = i - hw; j <= i + hw; j++) {
      if (j == i) continue;
      if (local != c.pts[ord[j]].p.cp_type) { local = 0; break; }
    }
    if (local != 0 && c.pts[ord[i]].p.cp_type != local) pending[ord[i]] = local;
  }
  for (const auto &kv : pending) c.pts[kv.first].p.cp_type = kv.second;
}

// feature_curve.hh:327-357
void smooth_interval_types(Curve &c) {
  const auto ord = to_ordinals(c);
  if (ord.empty()) return;
  const uint32_t front_type = c.pts[ord.front()].p.cp_type;
  for (int i = 0; i < ord.front(); i++) c.pts[i].p.cp_type = front_type;
  const uint32_t back_type = c.pts[ord.back()].p.cp_type;
  for (int i = ord.back(); i < (int)c.pts.size(); i++) c.pts[i].p.cp_type = back_type;
  for (int i = 0; i + 1 < (int)ord.size(); i++) {
    const uint32_t lt = c.pts[ord[i]].p.cp_type, rt = c.pts[ord[i + 1]].p.cp_type;
    if (lt == rt) {
      for (int j = ord[i]; j < ord[i + 1]; j++) c.pts[j].p.cp_type = lt;
    } else {   // the first point that changes type decides; the rest of the interval takes the right type
      int j;
      for (j = ord[i]; j < ord[i + 1]; j++)
        if (c.pts[j].p.cp_type != lt) break;
      for (; j < ord[i + 1]; j++) c.pts[j].p.cp_type = rt;
    }
  }
}

// feature_curve.hh:262-273
void rotate(Curve &c) {
  if (!c.loop || c.pts.empty()) return;
  if (c.pts.front().p.cp_type != c.pts.back().p.cp_type) return;
  size_t i;
  for (i = 0; i < c.pts.size(); i++)
    if (c.pts.front().p.cp_type != c.pts[i].p.cp_type) break;
  if (i < c.pts.size()) std::rotate(c.pts.begin(), c.pts.begin() + i, c.pts.end());
}

// feature_curve.hh:275-287
void reorder(Curve &c) {
  if (c.pts.empty() || c.loop) return;
  bool reverse = false;
  if (c.pts.front().p.timestep == c.pts.back().p.timestep) reverse = c.pts.front().p.t > c.pts.back().p.t;
  else reverse = c.pts.front().p.timestep > c.pts.back().p.timestep;
  if (reverse) std::reverse(c.pts.begin(), c.pts.end());
}

// feature_curve.hh:289-301
void adjust_time(Curve &c) {
  const size_t n = c.pts.size();
  for (size_t i = 0; i < n; i++) {
    if (i == 0 || c.pts[i].p.ordinal) continue;
    c.pts[i].p.t = std::max(c.pts[i - 1].p.t, c.pts[i].p.t);
  }
  for (size_t i = n; i-- > 0;) {
    if (i == n - 1 || c.pts[i].p.ordinal) continue;
    c.pts[i].p.t = std::min(c.pts[i + 1].p.t, c.pts[i].p.t);
  }
}

// feature_curve.hh:113-125 (discard: survivors into a fresh curve that keeps id / complete / loop; statistics refreshed)
template <typename F>
void discard(Curve &c, F drop) {
  Curve t;
  t.id = c.id; t.complete = c.complete; t.loop = c.loop;
  for (const auto &q : c.pts)
    if (!drop(q)) t.pts.push_back(q);
  update_statistics(t);
  c = t;
}

void discard_interval_points(Curve &c) { discard(c, [](const ftkb_curve_point &q) { return !q.p.ordinal; }); }
void discard_degenerate_points(Curve &c) { discard(c, [](const ftkb_curve_point &q) { return q.p.cp_type == 0 || q.p.cp_type == 1; }); }

// feature_curve.hh:404-417
void derive_velocity(Curve &c) {
  const int n = (int)c.pts.size();
  if (n < 2) return;
  for (int k = 0; k < 3; k++)
    for (int i = 0; i < n; i++) {
      if (i == 0) c.pts[i].v[k] = c.pts[i + 1].p.x[k] - c.pts[i].p.x[k];
      else if (i == n - 1) c.pts[i].v[k] = c.pts[i].p.x[k] - c.pts[i - 1].p.x[k];
      else c.pts[i].v[k] = 0.5 * (c.pts[i + 1].p.x[k] - c.pts[i - 1].p.x[k]);
    }
}

// feature_curve.hh:226-250: maximal runs of one type; the sub-curves are default curves (loop / complete cleared)
std::vector<Curve> split(const Curve &c) {
  std::vector<Curve> out;
  Curve sub;
  uint32_t current = 0;
  const size_t n = c.pts.size();
  for (size_t i = 0; i < n; i++) {
    if (sub.pts.empty()) current = c.pts[i].p.cp_type;
    if (c.pts[i].p.cp_type == current) sub.pts.push_back(c.pts[i]);
    if (c.pts[i].p.cp_type != current || i == n - 1) {
      if (!sub.pts.empty()) {
        update_statistics(sub);
        out.push_back(sub);
        sub.pts.clear();
      }
    }
  }
  return out;
}

}  // namespace

int ftkb_curveset::add(Curve c) {
  const int id = curves.empty() ? 0 : curves.back().id + 1;
  relabel(c, id);
  curves.push_back(std::move(c));
  return id;
}

void ftkb_curveset::add(Curve c, int label) {
  relabel(c, label);
  auto it = std::upper_bound(curves.begin(), curves.end(), label, [](int l, const Curve &x) { return l < x.id; });
  curves.insert(it, std::move(c));
}

// feature_curve_set.hh:514-532
static void split_all(ftkb_curveset &s) {
  std::vector<Curve> result, keep;
  for (auto &c : s.curves) {
    if (!c.consistent_type) {
      auto subs = split(c);
      result.insert(result.end(), subs.begin(), subs.end());
    } else keep.push_back(std::move(c));
  }
  s.curves = std::move(keep);
  for (auto &t : result) {
    const int label = t.pts[0].id;
    s.add(std::move(t), label);
  }
}

// feature_curve.hh:364-381 per curve (the int bounds compare against the double times), empty results dropped
void ftkb_curveset::intercept(int t0, int t1) {
  std::vector<Curve> out;
  for (const Curve &c : curves) {
    if (c.pts.empty()) continue;           // (the reference would read back() of an empty vector here)
    if (t0 > c.pts.back().p.t || t1 < c.pts.front().p.t) continue;
    Curve r;                                // a default curve: loop / complete cleared
    for (const auto &q : c.pts)
      if (q.p.t >= t0 && q.p.t <= t1) r.pts.push_back(q);
    if (r.pts.empty()) continue;
    relabel(r, c.id);
    update_statistics(r);
    out.push_back(std::move(r));
  }
  curves = std::move(out);
}

int ftkb_curveset::post_process(const std::string &ops) {
  size_t pos = 0;
  while (pos <= ops.size()) {
    const size_t comma = ops.find(',', pos);
    const std::string op = ops.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
    pos = comma == std::string::npos ? ops.size() + 1 : comma + 1;
    if (op.empty()) continue;
    auto foreach = [&](void (*f)(Curve &), bool stats = true) {
      for (auto &c : curves) { f(c); if (stats) update_statistics(c); }
    };
    if (op == "smooth_types") foreach([](Curve &c) { smooth_ordinal_types(c); smooth_interval_types(c); });
    else if (op == "rotate") foreach(rotate);
    else if (op == "split") split_all(*this);
    else if (op == "discard_interval_points") foreach(discard_interval_points);
    else if (op == "discard_degenerate_points") foreach(discard_degenerate_points);
    else if (op == "reorder") foreach(reorder);
    else if (op == "adjust_time") foreach(adjust_time);
    else if (op == "update_statistics") foreach([](Curve &) {});
    else if (op == "derive_velocity") foreach([](Curve &c) { discard_interval_points(c); derive_velocity(c); });
    else if (op.rfind("duration_pruning", 0) == 0) {
      // the post-processor parses but ignores this op (feature_curve_set_post_processor.hh:45-47); with a value it
      // acts like json_interface::post_process (json_interface.hh:770-776): curves shorter than the threshold go
      const size_t colon = op.find(':');
      if (colon != std::string::npos) {
        const double thr = std::atof(op.c_str() + colon + 1);
        if (thr > 0) curves.erase(std::remove_if(curves.begin(), curves.end(), [&](const Curve &c) { return c.tmax - c.tmin < thr; }), curves.end());
      }
    } else if (op.rfind("intercept:", 0) == 0) {
      // feature_curve_set_t::intercept(t0, t1) (feature_curve_set.hh:534-545) in place; curves are assumed reordered
      const size_t c1 = op.find(':'), c2 = op.find(':', c1 + 1);
      if (c2 == std::string::npos) { error = "post_process: intercept needs intercept:t0:t1"; return FTKB_ERR_INVALID; }
      intercept(std::atoi(op.c_str() + c1 + 1), std::atoi(op.c_str() + c2 + 1));
    } else if (op.rfind("legacy", 0) == 0) {
      // json_interface::post_process(): "legacy[:duration_threshold[:discard_interval_points[:derive_velocities]]]"
      double thr = 0; int disc = 0, vel = 0;
      { std::vector<std::string> f; size_t a = 0; while (a <= op.size()) { const size_t b = op.find(':', a); f.push_back(op.substr(a, b == std::string::npos ? std::string::npos : b - a)); a = b == std::string::npos ? op.size() + 1 : b + 1; }
        if (f.size() > 1) thr = std::atof(f[1].c_str());
        if (f.size() > 2) disc = std::atoi(f[2].c_str());
        if (f.size() > 3) vel = std::atoi(f[3].c_str()); }
      foreach([](Curve &c) { smooth_ordinal_types(c); smooth_interval_types(c); rotate(c); });
      if (thr > 0) curves.erase(std::remove_if(curves.begin(), curves.end(), [&](const Curve &c) { return c.tmax - c.tmin < thr; }), curves.end());
      split_all(*this);
      if (disc) foreach(discard_interval_points, false);
      foreach([](Curve &c) { reorder(c); adjust_time(c); });
      if (vel) foreach([](Curve &c) { discard_interval_points(c); derive_velocity(c); });
    } else {
      error = "post_process: unknown operation '" + op + "'";     // the reference calls fatal(FTK_ERR_UNKNOWN_OPTIONS)
      return FTKB_ERR_INVALID;
    }
  }
  return FTKB_OK;
}

// ---- C ABI --------------------------------------------------------------------------------------------------
extern "C" int ftkb_curveset_create(const ftkb_point *pts, uint64_t npts, const uint64_t *offsets, const uint64_t *point_idx,
                                    const uint8_t *loop, uint64_t ntraj, ftkb_curveset **out) {
  if (!out || (ntraj && (!offsets || !point_idx)) || (npts && !pts)) return FTKB_ERR_INVALID;
  *out = nullptr;
  ftkb_curveset *s = new ftkb_curveset();
  for (uint64_t i = 0; i < ntraj; i++) {
    Curve c;
    c.loop = loop ? loop[i] != 0 : false;
    for (uint64_t j = offsets[i]; j < offsets[i + 1]; j++) {
      if (point_idx[j] >= npts) { delete s; return FTKB_ERR_INVALID; }
      ftkb_curve_point q;
      std::memset(&q, 0, sizeof(q));
      q.p = pts[point_idx[j]];
      c.pts.push_back(q);
    }
    s->add(std::move(c));          // traced_critical_points.add(curves): ids 0 .. n-1 in trace order
  }
  for (auto &c : s->curves) update_statistics(c);     // finalize() ends with update_traj_statistics() (2d_regular.hh:224, 3d_regular.hh:122)
  *out = s;
  return FTKB_OK;
}

extern "C" void ftkb_curveset_destroy(ftkb_curveset *s) { delete s; }

extern "C" int ftkb_curveset_post_process(ftkb_curveset *s, const char *ops) {
  if (!s || !ops) return FTKB_ERR_INVALID;
  return s->post_process(ops);
}

extern "C" const char *ftkb_curveset_last_error(const ftkb_curveset *s) { return s ? s->error.c_str() : ""; }

extern "C" int ftkb_curveset_size(const ftkb_curveset *s, uint64_t *ncurves, uint64_t *npoints) {
  if (!s) return FTKB_ERR_INVALID;
  uint64_t np = 0;
  for (const auto &c : s->curves) np += c.pts.size();
  if (ncurves) *ncurves = s->curves.size();
  if (npoints) *npoints = np;
  return FTKB_OK;
}

// critical_point_tracker::slice_traced_critical_points (critical_point_tracker.hh:819-835): the ordinal points of one
// timestep, curves in the set's order, points in curve order
extern "C" int ftkb_curveset_slice(const ftkb_curveset *s, int32_t timestep, ftkb_curve_point *out, uint64_t cap, uint64_t *n) {
  if (!s || !n) return FTKB_ERR_INVALID;
  uint64_t k = 0;
  for (const auto &c : s->curves)
    for (const auto &q : c.pts)
      if (q.p.ordinal && q.p.timestep == timestep) {
        if (out && k < cap) out[k] = q;
        k++;
      }
  *n = k;
  return FTKB_OK;
}

extern "C" int ftkb_curveset_get(const ftkb_curveset *s, ftkb_curve_info *infos, ftkb_curve_point *pts) {
  if (!s) return FTKB_ERR_INVALID;
  uint64_t first = 0;
  for (size_t i = 0; i < s->curves.size(); i++) {
    const Curve &c = s->curves[i];
    if (infos) {
      ftkb_curve_info &o = infos[i];
      std::memset(&o, 0, sizeof(o));
      o.id = c.id; o.loop = c.loop; o.complete = c.complete; o.consistent_type = c.consistent_type;
      o.first = first; o.count = c.pts.size();
      o.tmin = c.tmin; o.tmax = c.tmax; o.smin = c.smin; o.smax = c.smax; o.persistence = c.persistence;
      o.vmmin = c.vmmin; o.vmmax = c.vmmax;
      for (int k = 0; k < 3; k++) { o.bbmin[k] = c.bbmin[k]; o.bbmax[k] = c.bbmax[k]; }
    }
    if (pts) std::copy(c.pts.begin(), c.pts.end(), pts + first);
    first += c.pts.size();
  }
  return FTKB_OK;
}
