// Several GPUs behind one tracker: ftkb_group (include/ftkb200.h).
//
// The reference plumbs a device list into every filter (filter.hh:47-51 set_device_ids) and decomposes the work
// with DIY/MPI (regular_tracker.hh:120-151).  Here ONE process drives N devices through N ordinary contexts:
//
//   time chunks   timesteps are cut into chunks of `chunk_timesteps`; chunk c = steps [cB, (c+1)B) is swept by its own
//                 context on device c mod N and owns every simplex whose corner time lies in the chunk (the face-ownership
//                 rule of SURVEY.md 8e).  It needs layers cB .. (c+1)B: the first layer of chunk c+1 is pushed to both
//                 contexts (halo = one layer).
//   workers       one host thread per device executes that device's commands (push / advance / update) in order, so the
//                 caller's thread only hands snapshots over; devices sweep their chunks concurrently.  Host snapshots are
//                 borrowed until the push returns, exactly as with a single context.
//   factor        the quantisation factor is a running quantity (critical_point_tracker.hh:850-864).  A chunk inherits the
//                 running minimum of min non-zero |v| from its predecessor.  It may start before the predecessor has finished
//                 as soon as the minimum known so far already saturates the factor (2^21: nothing later can change it);
//                 until then chunks run one after another.  Results therefore equal the one-device run bit for bit.
//   merge         ftkb_group_finalize collects the (sparse) punctured simplices of every chunk into one context on the first
//                 device and runs the union-find + ordering there: components that cross chunk boundaries are united by
//                 the same neighbour search that unites them inside a chunk.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ftkb200.h"

namespace {

struct Worker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv, cv_space;
  std::deque<std::function<void()>> q;
  bool stop = false;
  void start(int device) {
    th = std::thread([this, device] {
      ftkb_bind_thread_to_device(device);      // host-side staging of this device's snapshots stays in the socket next to it
      for (;;) {
        std::function<void()> f;
        {
          std::unique_lock<std::mutex> lk(m);
          cv.wait(lk, [this] { return stop || !q.empty(); });
          if (q.empty()) return;
          f = std::move(q.front());
          q.pop_front();
        }
        cv_space.notify_one();
        f();
      }
    });
  }
  void post(std::function<void()> f) {
    {
      std::unique_lock<std::mutex> lk(m);
      cv_space.wait(lk, [this] { return q.size() < 256; });      // the caller never runs unboundedly ahead of a device
      q.push_back(std::move(f));
    }
    cv.notify_one();
  }
  void join() {
    { std::lock_guard<std::mutex> lk(m); stop = true; }
    cv.notify_one();
    if (th.joinable()) th.join();
  }
};

// completion flag a caller can wait on
struct Done {
  std::mutex m;
  std::condition_variable cv;
  bool done = false;
  int rc = FTKB_OK;
  void set(int r) { { std::lock_guard<std::mutex> lk(m); rc = r; done = true; } cv.notify_all(); }
  int wait() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [this] { return done; }); return rc; }
};

struct Chunk {
  int index = 0, device = 0;
  ftkb_ctx *ctx = nullptr;               // created by the device's worker when the chunk may start
  bool finished = false;                 // every sweep of the chunk has been confirmed
  double resolution_out = DBL_MAX;       // running minimum after the chunk (valid once finished)
  std::vector<ftkb_point> points;
  ftkb_stats stats{};
};

int nbits_of(double res) {
  if (!(res > 0) || res >= DBL_MAX) return 8;
  const int nb = (int)std::ceil(std::log2(1.0 / res));
  return nb < 8 ? 8 : (nb > 21 ? 21 : nb);
}

}  // namespace

// spatial mode (chunk_timesteps == 0, 3D): device s holds planes [a0, a1] of every snapshot -- its corner planes plus the ghost
// planes its stencils reach (the reference's spatial decomposition, regular_tracker.hh:126-149)
struct Slab {
  int device = 0, a0 = 0, a1 = 0;
  ftkb_ctx *ctx = nullptr;
  std::vector<ftkb_point> points;
  ftkb_stats stats{};
};

struct ftkb_group {
  ftkb_config cfg{};
  std::vector<int> devices;
  int B = 8;
  bool spatial = false;
  std::vector<Slab> slabs;
  double running = DBL_MAX;                        // spatial mode: running minimum of min non-zero |v| over the whole array
  std::vector<std::unique_ptr<Worker>> workers;
  std::deque<std::unique_ptr<Chunk>> chunks;       // by index
  std::mutex m;                                    // protects chunk state shared between workers
  std::condition_variable cv;
  double saturating = 0.0;                         // a running minimum known to saturate the factor (0: not known yet)
  int npushed = 0, nadvanced = 0;
  std::atomic<int> failed{FTKB_OK};
  std::string error;
  ftkb_ctx *root = nullptr;

  Chunk *chunk(int c) {
    std::lock_guard<std::mutex> lk(m);
    while ((int)chunks.size() <= c) {
      std::unique_ptr<Chunk> ch(new Chunk());
      ch->index = (int)chunks.size();
      ch->device = devices[ch->index % devices.size()];
      chunks.push_back(std::move(ch));
    }
    return chunks[c].get();
  }
  void fail(int rc, const std::string &msg) {
    std::lock_guard<std::mutex> lk(m);
    if (failed.load() == FTKB_OK) { failed.store(rc); error = msg; }
    cv.notify_all();
  }
  // runs on the chunk's worker: wait until the chunk may start, then create its context
  int ensure_ctx(Chunk *ch) {
    if (ch->ctx) return FTKB_OK;
    double res_init = cfg.resolution_init;
    if (ch->index > 0) {
      std::unique_lock<std::mutex> lk(m);
      Chunk *prev = chunks[ch->index - 1].get();
      cv.wait(lk, [&] { return failed.load() != FTKB_OK || prev->finished || saturating > 0.0; });
      if (failed.load() != FTKB_OK) return failed.load();
      res_init = prev->finished ? prev->resolution_out : saturating;
      if (res_init >= DBL_MAX) res_init = 0.0;
    }
    ftkb_config c = cfg;
    c.device = ch->device;
    c.start_timestep = cfg.start_timestep + ch->index * B;
    c.resolution_init = res_init;
    const int rc = ftkb_create(&c, &ch->ctx);
    if (rc) fail(rc, std::string("chunk context: ") + ftkb_last_error(nullptr));
    return rc;
  }
  // after a sweep of a chunk: publish a saturating minimum as soon as one is known (lets the next chunk start early)
  void publish_resolution(Chunk *ch) {
    { std::lock_guard<std::mutex> lk(m); if (saturating > 0.0) return; }
    ftkb_stats st;
    if (ftkb_get_stats(ch->ctx, &st) != FTKB_OK) return;          // (confirms the chunk's sweeps: only until saturation is known)
    if (nbits_of(st.resolution) == 21) {
      { std::lock_guard<std::mutex> lk(m); saturating = st.resolution; }
      cv.notify_all();
    }
  }
  void finish_chunk(Chunk *ch) {
    ftkb_stats st{};
    int rc = ftkb_get_stats(ch->ctx, &st);
    uint64_t n = 0;
    if (!rc) rc = ftkb_num_points(ch->ctx, &n);
    if (!rc) { ch->points.resize(n); rc = ftkb_get_points(ch->ctx, ch->points.data(), n); }
    if (rc) fail(rc, std::string("chunk ") + std::to_string(ch->index) + ": " + ftkb_last_error(ch->ctx));
    ch->stats = st;
    ftkb_destroy(ch->ctx);            // its layers and buffers go back to the device; the points stay on the host
    ch->ctx = nullptr;
    { std::lock_guard<std::mutex> lk(m); ch->resolution_out = st.resolution; ch->finished = true; }
    cv.notify_all();
  }
};

extern "C" const char *ftkb_group_last_error(const ftkb_group *g) { return g ? g->error.c_str() : "null group"; }

extern "C" int ftkb_group_create(const ftkb_config *cfg, const int32_t *device_ids, int32_t n_devices, int32_t chunk_timesteps, ftkb_group **out) {
  if (!cfg || !out || !device_ids || n_devices < 1 || n_devices > 64 || chunk_timesteps < 0) return FTKB_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return FTKB_ERR_NO_DEVICE; }
  for (int i = 0; i < n_devices; i++)
    if (device_ids[i] < 0 || device_ids[i] >= ndev) return FTKB_ERR_NO_DEVICE;
  ftkb_group *g = new ftkb_group();
  g->cfg = *cfg;
  g->devices.assign(device_ids, device_ids + n_devices);
  g->B = chunk_timesteps;
  g->spatial = chunk_timesteps == 0;
  if (g->spatial) {
    // z-slabs: the domain's corner planes are split evenly; slab s also sweeps the corner plane above its last one (the flat
    // simplices of that plane are found twice and merged as one), and its array reaches two planes past every vertex it uses
    // (gradient of the Jacobian's neighbours); the first / last slab hold the array's end planes so that min non-zero |v|
    // still covers the whole array
    const int D = cfg->dims[2], glb = cfg->lb[2], gub = cfg->ub[2], nplanes = gub - glb + 1;
    if (cfg->nd != 3 || nplanes < 2 * n_devices) { delete g; return FTKB_ERR_INVALID; }
    g->running = cfg->resolution_init > 0 ? cfg->resolution_init : DBL_MAX;
    for (int s = 0; s < n_devices; s++) {
      Slab sl;
      sl.device = device_ids[s];
      const int c0 = glb + (int)((long long)nplanes * s / n_devices), c1 = glb + (int)((long long)nplanes * (s + 1) / n_devices) - 1;
      const int vtop = std::min(c1 + 1, gub);
      sl.a0 = s == 0 ? 0 : std::max(0, c0 - 2);
      sl.a1 = s == n_devices - 1 ? D - 1 : std::min(D - 1, vtop + 2);
      ftkb_config c = *cfg;
      c.device = sl.device;
      c.dims[2] = sl.a1 - sl.a0 + 1;
      c.lb[2] = c0 - sl.a0;
      c.ub[2] = vtop - sl.a0;
      c.slab_offset = sl.a0;
      c.slab_global_dim = D;
      c.slab_global_lb = glb;
      c.slab_global_ub = gub;
      const int rc = ftkb_create(&c, &sl.ctx);
      if (rc) {
        for (auto &x : g->slabs) ftkb_destroy(x.ctx);
        delete g;
        return rc;
      }
      g->slabs.push_back(sl);
    }
  }
  for (int i = 0; i < n_devices; i++) {
    g->workers.emplace_back(new Worker());
    g->workers.back()->start(device_ids[i]);
  }
  *out = g;
  return FTKB_OK;
}

extern "C" void ftkb_group_destroy(ftkb_group *g) {
  if (!g) return;
  g->fail(FTKB_ERR_INVALID, "group destroyed");            // releases workers that wait for a predecessor
  for (auto &w : g->workers) w->join();
  for (auto &ch : g->chunks) if (ch->ctx) ftkb_destroy(ch->ctx);
  for (auto &sl : g->slabs) if (sl.ctx) ftkb_destroy(sl.ctx);
  if (g->root) ftkb_destroy(g->root);
  delete g;
}

// ---- spatial mode: every device works on every step ------------------------------------------------------------------------
// run f(slab) on every device's worker and wait for all
static int on_all_slabs(ftkb_group *g, std::function<int(Slab &)> f, const char *what) {
  std::vector<std::shared_ptr<Done>> waits;
  for (size_t s = 0; s < g->slabs.size(); s++) {
    std::shared_ptr<Done> d(new Done());
    waits.push_back(d);
    Slab *sl = &g->slabs[s];
    g->workers[s]->post([g, sl, f, d, what] {
      int rc = g->failed.load();
      if (!rc) {
        rc = f(*sl);
        if (rc) g->fail(rc, std::string(what) + ": " + ftkb_last_error(sl->ctx));
      }
      d->set(rc);
    });
  }
  int rc = FTKB_OK;
  for (auto &d : waits) { const int r = d->wait(); if (r && !rc) rc = r; }
  return rc;
}

// The quantisation factor must be the undivided run's on every slab: until the running minimum saturates it, the slabs
// resolve the new layer first (a cells-only build for scalar input, a reduction for vector input), the minima are combined
// and handed back before anyone sweeps.  Once saturated (nbits = 21) nothing can change the factor and the step is skipped.
static int spatial_sync_resolution(ftkb_group *g) {
  if (nbits_of(g->running) == 21) return FTKB_OK;
  std::vector<double> res(g->slabs.size(), DBL_MAX);
  int rc = on_all_slabs(g, [&](Slab &sl) { return ftkb_last_layer_resolution(sl.ctx, &res[&sl - g->slabs.data()]); }, "resolution");
  if (rc) return rc;
  for (double r : res) if (r > 0 && r < g->running) g->running = r;
  if (g->running < DBL_MAX) {
    const double run = g->running;
    rc = on_all_slabs(g, [run](Slab &sl) { return ftkb_set_resolution(sl.ctx, run); }, "set_resolution");
  }
  return rc;
}

static int spatial_push(ftkb_group *g, std::function<int(Slab &)> push) {
  int rc = on_all_slabs(g, push, "push");
  if (!rc) rc = spatial_sync_resolution(g);
  g->npushed++;
  return rc;
}

// hand one snapshot to chunk c's context (runs on that device's worker)
static void post_push(ftkb_group *g, int c, std::function<int(ftkb_ctx *)> push, std::shared_ptr<Done> done) {
  Chunk *ch = g->chunk(c);
  Worker *w = g->workers[ch->index % g->workers.size()].get();
  w->post([g, ch, push, done] {
    int rc = g->failed.load();
    if (!rc) rc = g->ensure_ctx(ch);
    if (!rc) {
      rc = push(ch->ctx);
      if (rc) g->fail(rc, std::string("push: ") + ftkb_last_error(ch->ctx));
    }
    if (done) done->set(rc);
  });
}

static int push_layer(ftkb_group *g, std::function<int(ftkb_ctx *)> push, bool wait) {
  if (g->failed.load()) return g->failed.load();
  const int k = g->npushed++;
  const int c = k / g->B;
  std::vector<std::shared_ptr<Done>> waits;
  auto mk = [&]() -> std::shared_ptr<Done> { if (!wait) return nullptr; waits.emplace_back(new Done()); return waits.back(); };
  // the first layer of a chunk is also the last layer (halo) of the chunk before it
  if (k % g->B == 0 && k > 0) post_push(g, c - 1, push, mk());
  post_push(g, c, push, mk());
  int rc = FTKB_OK;
  for (auto &d : waits) { const int r = d->wait(); if (r && !rc) rc = r; }
  return rc;
}

// The first layer of chunk c > 0 arrives while chunk c-1 may still have to finish before chunk c's context can exist (it
// inherits the running minimum), and chunk c-1 finishes with an advance the caller only issues after this push has returned.
// So that layer is copied to the new chunk's device right away, into buffers of the group, and handed to the context later.
struct Staged { double *S = nullptr, *V = nullptr, *J = nullptr; };

extern "C" int ftkb_group_push_snapshot(ftkb_group *g, const double *scalar, const double *vector, const double *jacobian) {
  if (!g) return FTKB_ERR_INVALID;
  if (g->failed.load()) return g->failed.load();
  if (g->spatial) {
    // every device copies its planes (ghost planes included) over its own PCIe link, concurrently
    const size_t plane = (size_t)g->cfg.dims[0] * (size_t)g->cfg.dims[1];
    return spatial_push(g, [=](Slab &sl) {
      const size_t o = plane * (size_t)sl.a0;
      return ftkb_push_snapshot(sl.ctx, scalar ? scalar + o : nullptr, vector ? vector + 3 * o : nullptr, jacobian ? jacobian + 9 * o : nullptr, FTKB_MEM_HOST);
    });
  }
  const int k = g->npushed;
  if (k % g->B != 0 || k == 0)
    // host memory, borrowed until return: wait for the copy
    return push_layer(g, [=](ftkb_ctx *c) { return ftkb_push_snapshot(c, scalar, vector, jacobian, FTKB_MEM_HOST); }, true);
  g->npushed++;
  const int c = k / g->B;
  size_t nvert = 1;
  for (int j = 0; j < g->cfg.nd; j++) nvert *= (size_t)g->cfg.dims[j];
  const size_t n = (size_t)g->cfg.nd;
  std::shared_ptr<Done> d_halo(new Done()), d_stage(new Done());
  post_push(g, c - 1, [=](ftkb_ctx *x) { return ftkb_push_snapshot(x, scalar, vector, jacobian, FTKB_MEM_HOST); }, d_halo);   // halo of the chunk before
  Chunk *ch = g->chunk(c);
  Worker *w = g->workers[ch->index % g->workers.size()].get();
  std::shared_ptr<Staged> st(new Staged());
  w->post([=] {
    int rc = g->failed.load();
    auto up = [&](const double *src, size_t count, double **dst) {
      if (!src || rc) return;
      if (cudaMalloc(dst, count * sizeof(double)) != cudaSuccess || cudaMemcpy(*dst, src, count * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        rc = FTKB_ERR_NOMEM;
        g->fail(rc, "group: staging the first layer of a chunk failed");
      }
    };
    if (!rc && cudaSetDevice(ch->device) != cudaSuccess) { cudaGetLastError(); rc = FTKB_ERR_CUDA; g->fail(rc, "cudaSetDevice"); }
    up(scalar, nvert, &st->S); up(vector, nvert * n, &st->V); up(jacobian, nvert * n * n, &st->J);
    d_stage->set(rc);
  });
  w->post([=] {
    int rc = g->failed.load();
    if (!rc) rc = g->ensure_ctx(ch);
    if (!rc) {
      rc = ftkb_push_snapshot(ch->ctx, st->S, st->V, st->J, FTKB_MEM_DEVICE);      // copied (and synchronised) before it returns
      if (rc) g->fail(rc, std::string("push: ") + ftkb_last_error(ch->ctx));
    }
    cudaSetDevice(ch->device);
    cudaFree(st->S); cudaFree(st->V); cudaFree(st->J);
  });
  const int r1 = d_halo->wait(), r2 = d_stage->wait();
  return r1 ? r1 : r2;
}

extern "C" int ftkb_group_push_synthetic(ftkb_group *g, int kind, const double *params, int nparams, double t) {
  if (!g || nparams < 0 || nparams > 8 || (nparams && !params)) return FTKB_ERR_INVALID;
  std::vector<double> p(params, params + nparams);
  if (g->spatial) return spatial_push(g, [=](Slab &sl) { return ftkb_push_synthetic(sl.ctx, kind, p.data(), (int)p.size(), t); });
  return push_layer(g, [=](ftkb_ctx *c) { return ftkb_push_synthetic(c, kind, p.data(), (int)p.size(), t); }, false);
}

static int post_sweep(ftkb_group *g, int step, bool advance) {
  if (g->failed.load()) return g->failed.load();
  const int c = step / g->B;
  Chunk *ch = g->chunk(c);
  Worker *w = g->workers[ch->index % g->workers.size()].get();
  const bool last_of_chunk = advance && (step + 1) % g->B == 0;
  w->post([g, ch, advance, last_of_chunk] {
    int rc = g->failed.load();
    if (!rc) rc = g->ensure_ctx(ch);
    if (rc) return;
    rc = advance ? ftkb_advance_timestep(ch->ctx) : ftkb_update_timestep(ch->ctx);
    if (rc) { g->fail(rc, std::string("sweep: ") + ftkb_last_error(ch->ctx)); return; }
    g->publish_resolution(ch);
    if (last_of_chunk || !advance) g->finish_chunk(ch);
  });
  return FTKB_OK;
}

extern "C" int ftkb_group_advance_timestep(ftkb_group *g) {
  if (!g) return FTKB_ERR_INVALID;
  if (g->npushed < g->nadvanced + 2) { g->error = "advance_timestep: two snapshots are needed"; return FTKB_ERR_INVALID; }
  if (g->spatial) { g->nadvanced++; return on_all_slabs(g, [](Slab &sl) { return ftkb_advance_timestep(sl.ctx); }, "sweep"); }
  return post_sweep(g, g->nadvanced++, true);
}

extern "C" int ftkb_group_update_timestep(ftkb_group *g) {
  if (!g) return FTKB_ERR_INVALID;
  if (g->npushed < g->nadvanced + 1) { g->error = "update_timestep: no snapshot has been pushed"; return FTKB_ERR_INVALID; }
  if (g->spatial) return on_all_slabs(g, [](Slab &sl) { return ftkb_update_timestep(sl.ctx); }, "sweep");
  return post_sweep(g, g->nadvanced, false);
}

extern "C" int ftkb_group_finalize(ftkb_group *g, ftkb_ctx **root) {
  if (!g || !root) return FTKB_ERR_INVALID;
  *root = nullptr;
  if (g->spatial) {
    int rc = on_all_slabs(g, [](Slab &sl) {
      uint64_t n = 0;
      int r = ftkb_get_stats(sl.ctx, &sl.stats);
      if (!r) r = ftkb_num_points(sl.ctx, &n);
      if (!r) { sl.points.resize(n); r = ftkb_get_points(sl.ctx, sl.points.data(), n); }
      return r;
    }, "collect");
    if (rc) return rc;
    if (g->root) { ftkb_destroy(g->root); g->root = nullptr; }
    ftkb_config c = g->cfg;
    c.device = g->devices[0];
    if ((rc = ftkb_create(&c, &g->root))) { g->error = ftkb_last_error(nullptr); return rc; }
    // (the flat simplices of a cut plane were found by both neighbours: the root's sort drops the duplicates)
    for (auto &sl : g->slabs)
      if (!sl.points.empty() && (rc = ftkb_import_points(g->root, sl.points.data(), sl.points.size()))) { g->error = ftkb_last_error(g->root); return rc; }
    if ((rc = ftkb_finalize(g->root))) { g->error = ftkb_last_error(g->root); return rc; }
    *root = g->root;
    return FTKB_OK;
  }
  // wait for every device's queue to drain
  std::vector<std::shared_ptr<Done>> waits;
  for (auto &w : g->workers) {
    std::shared_ptr<Done> d(new Done());
    waits.push_back(d);
    w->post([d] { d->set(FTKB_OK); });
  }
  for (auto &d : waits) d->wait();
  if (g->failed.load()) return g->failed.load();
  // chunks that never reached their last sweep (the series ended inside them, without a final update): collect them as they are
  for (auto &ch : g->chunks)
    if (ch->ctx && !ch->finished) g->finish_chunk(ch.get());
  if (g->failed.load()) return g->failed.load();
  if (g->root) { ftkb_destroy(g->root); g->root = nullptr; }
  ftkb_config c = g->cfg;
  c.device = g->devices[0];
  int rc = ftkb_create(&c, &g->root);
  if (rc) { g->error = ftkb_last_error(nullptr); return rc; }
  for (auto &ch : g->chunks)
    if (!ch->points.empty() && (rc = ftkb_import_points(g->root, ch->points.data(), ch->points.size()))) { g->error = ftkb_last_error(g->root); return rc; }
  if ((rc = ftkb_finalize(g->root))) { g->error = ftkb_last_error(g->root); return rc; }
  *root = g->root;
  return FTKB_OK;
}

extern "C" int ftkb_group_get_stats(ftkb_group *g, ftkb_stats *sum, int32_t *chunks_done) {
  if (!g || !sum) return FTKB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(g->m);
  ftkb_stats s{};
  int done = 0;
  for (auto &sl : g->slabs) {          // spatial mode: as collected by ftkb_group_finalize
    const ftkb_stats &t = sl.stats;
    s.simplices_tested += t.simplices_tested; s.cells_scanned += t.cells_scanned; s.cells_refined += t.cells_refined;
    s.points += t.points; s.kernel_launches += t.kernel_launches; s.h2d_bytes += t.h2d_bytes; s.d2h_bytes += t.d2h_bytes;
    s.ms_derive += t.ms_derive; s.ms_scan += t.ms_scan; s.ms_test += t.ms_test;
    s.scan_launches += t.scan_launches; s.sweeps_repeated += t.sweeps_repeated;
    s.scaling_factor = t.scaling_factor; s.resolution = t.resolution;
    done++;
  }
  for (auto &ch : g->chunks) {
    if (!ch->finished) continue;
    done++;
    const ftkb_stats &t = ch->stats;
    s.simplices_tested += t.simplices_tested; s.cells_scanned += t.cells_scanned; s.cells_refined += t.cells_refined;
    s.points += t.points; s.kernel_launches += t.kernel_launches; s.h2d_bytes += t.h2d_bytes; s.d2h_bytes += t.d2h_bytes;
    s.ms_derive += t.ms_derive; s.ms_scan += t.ms_scan; s.ms_test += t.ms_test;
    s.scan_launches += t.scan_launches; s.sweeps_repeated += t.sweeps_repeated;
    s.scaling_factor = t.scaling_factor; s.resolution = t.resolution;
  }
  *sum = s;
  if (chunks_done) *chunks_done = done;
  return FTKB_OK;
}
