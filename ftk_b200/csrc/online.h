// Streaming trajectories: the reference's trace_critical_points_online (critical_point_tracker.hh:522-641),
// the grow step run after every interval sweep when set_enable_streaming_trajectories(true).
#pragma once
#include <cstdint>
#include <deque>
#include <vector>

#include "../../include/ftkb200.h"
#include "mesh_tables.h"

namespace ftkb {

struct OnlineCurve {
  std::deque<ftkb_point> pts;
  bool loop = false, complete = false;
  OnlineCurve() = default;
  OnlineCurve(const OnlineCurve &) = default;
  OnlineCurve &operator=(const OnlineCurve &) = default;
  // noexcept so that a growing std::vector<OnlineCurve> moves its curves instead of copying every point
  OnlineCurve(OnlineCurve &&o) noexcept : pts(std::move(o.pts)), loop(o.loop), complete(o.complete) {}
  OnlineCurve &operator=(OnlineCurve &&o) noexcept { pts = std::move(o.pts); loop = o.loop; complete = o.complete; return *this; }
};

class OnlineTracer {
 public:
  OnlineTracer(int nd, const int32_t lb[3], const int32_t ub[3]);
  // one grow() = one call of trace_critical_points_online on the punctured simplices found since the last one
  void grow(const ftkb_point *pts, uint64_t n);
  const std::vector<OnlineCurve> &curves() const { return curves_; }
  uint64_t npoints() const;
  // element key of a point (the reference's element order packed into 64 bits; same packing as the device sort)
  bool key_of(const ftkb_point &p, uint64_t &key) const { return key_at(p.corner[0], p.corner[1], p.corner[2], p.corner[3], p.simplex_type, key); }

 private:
  bool key_at(int x, int y, int z, int t, int type, uint64_t &key) const;
  int neighbor_keys(const ftkb_point &p, uint64_t out[9]) const;   // ascending, the element itself included

  int nd_;
  int32_t lb_[3], ub_[3];
  int64_t ny_, nz_;
  // neighbour candidates of every simplex type (the element itself included) in ascending element order: the order
  // of (corner + offset, type) does not depend on the corner
  struct Candidate { int8_t off[4]; int8_t type; int64_t cell_delta; };
  Candidate cand_[60][9];
  int ncand_[60] = {};
  int ntypes_ = 0;             // 12 (2D+t) or 60 (3D+t); points with another simplex type are ignored
  std::vector<OnlineCurve> curves_;
};

}  // namespace ftkb
