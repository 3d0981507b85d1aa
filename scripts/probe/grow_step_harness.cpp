#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/ftkb200.h"
int main(int argc, char **argv) {
  FILE *f = fopen(argv[1], "rb");
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<char> raw(sz); fread(raw.data(), 1, sz, f); fclose(f);
  // npy header: 10 bytes + header_len
  unsigned short hl = *(unsigned short *)(raw.data() + 8);
  const char *data = raw.data() + 10 + hl;
  size_t n = (sz - 10 - hl) / sizeof(ftkb_point);
  const ftkb_point *p = (const ftkb_point *)data;
  int W = atoi(argv[2]), T = atoi(argv[3]);
  std::vector<std::vector<ftkb_point>> steps(T);
  for (size_t i = 0; i < n; i++) steps[p[i].corner[3]].push_back(p[i]);
  double best = 1e9;
  for (int r = 0; r < 7; r++) {
    int32_t lb[3] = {2, 2, 0}, ub[3] = {W - 2, W - 2, 0};
    ftkb_online *o; ftkb_online_create(2, lb, ub, &o);
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < T; t++) ftkb_online_grow(o, steps[t].data(), steps[t].size());
    double dt = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (dt < best) best = dt;
    uint64_t nt, np; ftkb_online_size(o, &nt, &np);
    if (r == 0) printf("%zu points, %llu trajectories, %llu kept\n", n, (unsigned long long)nt, (unsigned long long)np);
    ftkb_online_destroy(o);
  }
  printf("grow %.3f ms = %.0f ns/point\n", best, best * 1e6 / n);
}
