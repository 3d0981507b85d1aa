// Host-side curve sets: the traced trajectories as mutable curves, and the reference's trajectory
// post-processing on them.  No device code; the C ABI exposes it both for a finalized
// context and for caller-provided trajectories (so it can be tested without a GPU).
//
// ref: include/ftk/features/feature_curve.hh (per-curve operations), feature_curve_set.hh:447-532 (multimap
// keyed by curve id: add / filter / split_all), filters/feature_curve_set_post_processor.hh:23-70 (op names),
// filters/json_interface.hh:758-800 (the legacy post_process() sequence).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/ftkb200.h"

struct ftkb_curveset {
  struct Curve {
    int32_t id = 0;
    bool loop = false, complete = false;
    uint32_t consistent_type = 0;          // 0 unless update_statistics() found one type (feature_curve.hh:52)
    double tmin = 0, tmax = 0, bbmin[3] = {0, 0, 0}, bbmax[3] = {0, 0, 0};
    double smin = 0, smax = 0, persistence = 0, vmmin = 0, vmmax = 0;
    std::vector<ftkb_curve_point> pts;
  };
  // multimap<int, curve> order: ascending id, equal ids in insertion order
  std::vector<Curve> curves;
  std::string error;

  int add(Curve c);                        // feature_curve_set.hh:458-465: id = last id + 1, points relabelled
  void add(Curve c, int label);            // :467-472
  void intercept(int t0, int t1);         // feature_curve_set.hh:534-545
  int post_process(const std::string &ops);
};
