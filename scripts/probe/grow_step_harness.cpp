// Host-only timing of the streaming grow step: reads the .npy of ftkb_point records scripts/grow_step_timing.py caches
// (/tmp/grow_pts_W_T.npy), optionally tiles the field R x R times, and replays it one timestep at a time.
//   g++ -O2 -std=c++17 -I ftk_b200/csrc -I include -I /usr/local/cuda/include scripts/probe/grow_step_harness.cpp \
//       ftk_b200/csrc/online.cpp ftk_b200/csrc/mesh_tables.cpp -o /tmp/grow_harness
//   /tmp/grow_harness /tmp/grow_pts_768_16.npy 768 16 [prepared=0|1] [R]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/ftkb200.h"
int main(int argc, char **argv) {
  if (argc < 4) return 1;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 1;
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<char> raw(sz);
  if (fread(raw.data(), 1, sz, f) != (size_t)sz) return 1;
  fclose(f);
  const unsigned short hl = *(unsigned short *)(raw.data() + 8);      // npy v1 header: 10 bytes + header_len
  const char *data = raw.data() + 10 + hl;
  const size_t n = (sz - 10 - hl) / sizeof(ftkb_point);
  const ftkb_point *p = (const ftkb_point *)data;
  const int W = atoi(argv[2]), T = atoi(argv[3]);
  const bool prepared = argc > 4 && atoi(argv[4]) != 0;
  const int R = argc > 5 ? atoi(argv[5]) : 1;
  std::vector<std::vector<ftkb_point>> steps(T);
  for (int a = 0; a < R; a++)
    for (int b = 0; b < R; b++)
      for (size_t i = 0; i < n; i++) {
        ftkb_point q = p[i];
        q.corner[0] += a * W; q.corner[1] += b * W;
        steps[q.corner[3]].push_back(q);
      }
  double best = 1e9;
  for (int r = 0; r < 7; r++) {
    int32_t lb[3] = {2, 2, 0}, ub[3] = {W * R - 2, W * R - 2, 0};
    ftkb_online *o;
    ftkb_online_create(2, lb, ub, &o);
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < T; t++) (prepared ? ftkb_online_grow_prepared : ftkb_online_grow)(o, steps[t].data(), steps[t].size());
    const double dt = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (dt < best) best = dt;
    uint64_t nt, np;
    ftkb_online_size(o, &nt, &np);
    if (r == 0) printf("%zu points, %llu trajectories, %llu on trajectories\n", n * R * R, (unsigned long long)nt, (unsigned long long)np);
    ftkb_online_destroy(o);
  }
  printf("grow (%s) %.3f ms = %.0f ns/point\n", prepared ? "prepared on the host + index-list walk" : "hash of the batch", best, best * 1e6 / (n * R * R));
}
