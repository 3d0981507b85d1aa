// ftkb200 -- command-line front end with the `ftk -f cp` flags of the reference
// (ref: src/cli/ftk.cpp:885-1078 option table and main; include/ftk/filters/json_interface.hh:606-725
// consume_regular: the push / advance / update loop; include/ftk/ndarray/stream.hh:258-440,1443-1568
// synthetic stream defaults and per-timestep time conventions).
//
// Host code only: inputs are produced (closed-form generators, same formulas and libm as the
// reference, so snapshots are bit-identical) or read (raw float32/float64 series), pushed through
// the C++ tracker classes of include/ftk_b200/critical_point_tracker_regular.hh, which call the
// C ABI of libftkb200.so; every per-simplex computation runs on the GPU.  Without a B200 the
// program fails with the library's error message -- there is no CPU path.
#include <algorithm>
#include <chrono>
#include <fcntl.h>
#include <unistd.h>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ftk_b200/critical_point_tracker_regular.hh"

using namespace ftk_b200;

namespace {

struct Options {
  std::string feature = "cp", input, input_format = "float64", synthetic, output, output_type = "traced", output_format = "auto";
  std::string type_filter, accelerator = "cuda", var, post_process;
  long width = -1, height = -1, depth = -1, timesteps = -1;
  int device = 0, nthreads = 0, time_chunk = 8;
  std::vector<int> devices;
  bool timing = false, compute_degrees = false, no_robust = false, verbose = false, device_generators = false, help = false, stream = false;
  std::vector<double> x0, dir;
  double time_scale = 0.1;
};

[[noreturn]] void die(const std::string &m) {
  std::fprintf(stderr, "ftkb200: %s\n", m.c_str());
  std::exit(1);
}

void usage() {
  std::puts(
      "usage: ftkb200 -f cp (--synthetic NAME | --input PATTERN) -o OUTPUT [options]\n"
      "  -f, --feature cp|critical_point   feature type (only critical points are implemented)\n"
      "      --synthetic NAME              woven | moving_extremum_2d | moving_extremum_3d | double_gyre | merger_2d | tornado\n"
      "  -i, --input PATTERN               raw file: one file holding all timesteps, or a printf pattern (e.g. s-%03d.raw)\n"
      "      --input-format float32|float64\n"
      "      --var a[,b[,c]]               variable names; 2 / 3 names = vector field with the component index fastest\n"
      "  -w, --width N  -h, --height N  -d, --depth N  -n, --timesteps N\n"
      "      --x0 a,b[,c]  --dir a,b[,c]   moving_extremum parameters;  --time-scale s (double_gyre)\n"
      "  -o, --output FILE                 result file\n"
      "      --output-type traced|discrete|sliced  (default traced; sliced: OUTPUT is a printf pattern, one text file per timestep)\n"
      "      --output-format text|json|binary  (default: by file name, .json -> json, else text; binary = the reference's DIY archive)\n"
      "      --type-filter min|max|saddle|...  (2D; names joined with |)\n"
      "      --post-process OPS                smooth_types,rotate,split,discard_interval_points,reorder,adjust_time,derive_velocity,...\n"
      "      --compute-degrees  --no-robust-detection  --timing  -v/--verbose\n"
      "      --stream                          streaming trajectories: grown after every timestep (trace_critical_points_online)\n"
      "  -a, --accelerator cuda            (the only back end)   --device ID[,ID...]   --time-chunk N (timesteps per device turn, default 8; 0 = z-slabs, 3D)\n"
      "      --nthreads N                  (ignored)\n"
      "      --device-generators           synthesise inputs on the GPU (CUDA libm; not bit-identical to the host generators)");
}

std::vector<double> parse_list(const std::string &s) {
  std::vector<double> v;
  size_t a = 0;
  while (a <= s.size()) {
    const size_t b = s.find(',', a);
    v.push_back(std::atof(s.substr(a, b == std::string::npos ? std::string::npos : b - a).c_str()));
    if (b == std::string::npos) break;
    a = b + 1;
  }
  return v;
}

Options parse(int argc, char **argv) {
  Options o;
  auto need = [&](int &i) -> std::string { if (i + 1 >= argc) die(std::string("missing value for ") + argv[i]); return argv[++i]; };
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "-f" || a == "--feature") o.feature = need(i);
    else if (a == "-i" || a == "--input") o.input = need(i);
    else if (a == "--input-format") o.input_format = need(i);
    else if (a == "--synthetic") o.synthetic = need(i);
    else if (a == "-w" || a == "--width") o.width = std::atol(need(i).c_str());
    else if (a == "-h" || a == "--height") o.height = std::atol(need(i).c_str());
    else if (a == "-d" || a == "--depth") o.depth = std::atol(need(i).c_str());
    else if (a == "-n" || a == "--timesteps") o.timesteps = std::atol(need(i).c_str());
    else if (a == "--var") o.var = need(i);
    else if (a == "--x0") o.x0 = parse_list(need(i));
    else if (a == "--dir") o.dir = parse_list(need(i));
    else if (a == "--time-scale") o.time_scale = std::atof(need(i).c_str());
    else if (a == "-o" || a == "--output") o.output = need(i);
    else if (a == "--output-type") o.output_type = need(i);
    else if (a == "--output-format") o.output_format = need(i);
    else if (a == "--type-filter") o.type_filter = need(i);
    else if (a == "--nthreads") o.nthreads = std::atoi(need(i).c_str());
    else if (a == "--thread-backend" || a == "--device-buffer" || a == "--nblocks") need(i);   // accepted, meaningless here
    else if (a == "--affinity" || a == "--async") {}
    else if (a == "--timing") o.timing = true;
    else if (a == "-a" || a == "--accelerator") o.accelerator = need(i);
    else if (a == "--device") {        // one id, or a comma-separated list: the tracker then runs on all of them (filter.hh:47-51)
      o.devices.clear();
      std::string v = need(i);
      for (size_t pos = 0; pos <= v.size();) {
        const size_t q = v.find(',', pos);
        const std::string tok = v.substr(pos, q == std::string::npos ? std::string::npos : q - pos);
        if (!tok.empty()) o.devices.push_back(std::atoi(tok.c_str()));
        if (q == std::string::npos) break;
        pos = q + 1;
      }
      if (o.devices.empty()) die("--device needs an id or a list of ids");
      o.device = o.devices[0];
    }
    else if (a == "--time-chunk") o.time_chunk = std::atoi(need(i).c_str());
    else if (a == "--compute-degrees") o.compute_degrees = true;
    else if (a == "--no-robust-detection") o.no_robust = true;
    else if (a == "--device-generators") o.device_generators = true;
    else if (a == "-v" || a == "--verbose") o.verbose = true;
    else if (a == "--help") o.help = true;
    else if (a == "--post-process") o.post_process = need(i);      // src/cli/ftk.cpp:253-257, :971
    else if (a == "--stream") o.stream = true;                      // src/cli/ftk.cpp:47,202-203: enable_streaming_trajectories
    else die("unknown option " + a);
  }
  return o;
}

// ref: src/cli/ftk.cpp:104-117 (names joined with '|')
unsigned parse_type_filter(const std::string &s) {
  unsigned f = 0;
  size_t a = 0;
  while (a <= s.size()) {
    const size_t b = s.find('|', a);
    const std::string w = s.substr(a, b == std::string::npos ? std::string::npos : b - a);
    if (w == "degenerate") f |= CRITICAL_POINT_2D_DEGENERATE;
    else if (w == "min" || w == "repelling") f |= CRITICAL_POINT_2D_MINIMUM;
    else if (w == "max" || w == "attracting") f |= CRITICAL_POINT_2D_MAXIMUM;
    else if (w == "saddle") f |= CRITICAL_POINT_2D_SADDLE;
    else if (w == "attracting_focus") f |= CRITICAL_POINT_2D_ATTRACTING_FOCUS;
    else if (w == "repelling_focus") f |= CRITICAL_POINT_2D_REPELLING_FOCUS;
    else if (w == "center") f |= CRITICAL_POINT_2D_CENTER;
    else die("unknown critical point type '" + w + "'");
    if (b == std::string::npos) break;
    a = b + 1;
  }
  return f;
}

// ---- host generators: closed forms of include/ftk/ndarray/synthetic.hh, evaluated with the host libm -------------
void gen_woven(long W, long H, double t, double *out) {               // synthetic.hh:32-48 (scaling factor 15)
  for (long j = 0; j < H; j++)
    for (long i = 0; i < W; i++) {
      const double x = ((double(i) / (W - 1)) - 0.5) * 15.0, y = ((double(j) / (H - 1)) - 0.5) * 15.0;
      out[i + W * j] = std::cos(x * std::cos(t) - y * std::sin(t)) * std::sin(x * std::sin(t) + y * std::cos(t));
    }
}

void gen_merger(long W, long H, double t, double *out) {              // synthetic.hh:262-297
  const double cx0 = std::sin(t - M_PI_2), cx1 = std::sin(t + M_PI_2), cy = 1e-4;
  for (long j = 0; j < H; j++)
    for (long i = 0; i < W; i++) {
      double x = ((double(i) / (W - 1)) - 0.5) * 4.0, y = ((double(j) / (H - 1)) - 0.5) * 4.0;
      const double xp = x * std::cos(t) - y * std::sin(t), yp = x * std::sin(t) + y * std::cos(t);
      x = xp; y = yp;
      const double f0 = std::exp(-((x - cx0) * (x - cx0) + (y - cy) * (y - cy))), f1 = std::exp(-((x - cx1) * (x - cx1) + (y - cy) * (y - cy)));
      out[i + W * j] = std::max(f0, f1);
    }
}

void gen_moving_extremum(int nd, const long *dims, const double *x0, const double *dir, double t, double *out) {   // synthetic.hh:332-354
  const long W = dims[0], H = dims[1], D = nd == 3 ? dims[2] : 1;
  double xc[3] = {0, 0, 0};
  for (int q = 0; q < nd; q++) xc[q] = x0[q] + dir[q] * t;
  for (long k = 0; k < D; k++)
    for (long j = 0; j < H; j++)
      for (long i = 0; i < W; i++) {
        const double x[3] = {double(i), double(j), double(k)};
        double d = 0;
        for (int q = 0; q < nd; q++) d += std::pow(x[q] - xc[q], 2.0);
        out[i + W * (j + H * k)] = d;
      }
}

void gen_tornado(long xs, long ys, long zs, int time, double *out) {      // synthetic.hh:441-494
  const double SMALL = 0.00000000001;
  const double xdelta = 1.0 / (xs - 1.0), ydelta = 1.0 / (ys - 1.0), zdelta = 1.0 / (zs - 1.0);
  for (long iz = 0; iz < zs; iz++) {
    const double z = iz * zdelta;
    const double xc = 0.5 + 0.1 * std::sin(0.04 * time + 10.0 * z), yc = 0.5 + 0.1 * std::cos(0.03 * time + 3.0 * z);
    const double r = 0.1 + 0.4 * z * z + 0.1 * z * std::sin(8.0 * z), r2 = 0.2 + 0.1 * z;
    for (long iy = 0; iy < ys; iy++) {
      const double y = iy * ydelta;
      for (long ix = 0; ix < xs; ix++) {
        const double x = ix * xdelta;
        double temp = std::sqrt((y - yc) * (y - yc) + (x - xc) * (x - xc));
        double scale = std::fabs(r - temp);
        scale = scale > r2 ? 0.8 - scale : 1.0;
        double z0 = 0.1 * (0.1 - temp * z);
        if (z0 < 0.0) z0 = 0.0;
        temp = std::sqrt(temp * temp + z0 * z0);
        scale = (r + r2 - temp) * scale / (temp + SMALL);
        scale = scale / (1 + z);
        *out++ = scale * (y - yc) + 0.1 * (x - xc);
        *out++ = scale * -(x - xc) + 0.1 * (y - yc);
        *out++ = scale * z0;
      }
    }
  }
}

void gen_double_gyre(long W, long H, double time, double *out) {      // synthetic.hh:130-150,193-217: A = 0.1, omega = 2 pi, eps = 0.25
  const double A = 0.1, omega = M_PI * 2, eps = 0.25;
  for (long j = 0; j < H; j++)
    for (long i = 0; i < W; i++) {
      const double x = (double(i) / (W - 1)) * 2, y = double(j) / (H - 1);
      const double a = eps * std::sin(omega * time), b = 1 - 2 * eps * std::sin(omega * time);
      const double f = a * x * x + b * x, dfdx = 2 * a * x + b;
      out[2 * (i + W * j)] = -M_PI * A * std::sin(M_PI * f) * std::cos(M_PI * y);
      out[2 * (i + W * j) + 1] = M_PI * A * std::cos(M_PI * f) * std::sin(M_PI * y) * dfdx;
    }
}

// raw input: snapshot k from one big file (offset k * bytes) or from sprintf(pattern, k).  The byte range is read by several
// threads at once (pread on disjoint slices, float32 converted slice by slice): one thread copying out of the page cache
// delivers about 5 GB/s, a fifth of what the PCIe link behind it takes.
// (keep_f32: the float32 bytes are delivered as they are -- `out` then holds `count` floats -- and widened on the device)
void read_raw(const Options &o, long k, size_t count, bool f32, double *out, bool keep_f32 = false) {
  std::string path = o.input;
  long offset = 0;
  if (o.input.find('%') != std::string::npos) {
    char buf[4096];
    std::snprintf(buf, sizeof(buf), o.input.c_str(), (int)k);
    path = buf;
  } else {
    offset = k * (long)count * (f32 ? 4 : 8);
  }
  const int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) die("cannot open " + path);
  const size_t esz = f32 ? 4 : 8;
  const unsigned hw = std::thread::hardware_concurrency();
  const size_t nthr = std::max<size_t>(1, std::min<size_t>({(size_t)8, (size_t)(hw ? hw : 1), count / (1u << 20) + 1}));
  std::vector<std::thread> pool;
  std::vector<int> bad(nthr, 0);
  for (size_t t = 0; t < nthr; t++)
    pool.emplace_back([&, t] {
      const size_t e0 = count * t / nthr, e1 = count * (t + 1) / nthr;
      std::vector<float> tmp;
      const size_t chunk = 1u << 22;                       // elements per pread
      if (f32 && !keep_f32) tmp.resize(std::min(chunk, e1 - e0));
      for (size_t e = e0; e < e1;) {
        const size_t n = std::min(chunk, e1 - e);
        char *dst = keep_f32 ? reinterpret_cast<char *>(reinterpret_cast<float *>(out) + e)
                             : f32 ? reinterpret_cast<char *>(tmp.data()) : reinterpret_cast<char *>(out + e);
        size_t got = 0;
        while (got < n * esz) {
          const ssize_t r = ::pread(fd, dst + got, n * esz - got, (off_t)(offset + (long)(e * esz + got)));
          if (r <= 0) { bad[t] = 1; return; }
          got += (size_t)r;
        }
        if (f32 && !keep_f32) for (size_t i = 0; i < n; i++) out[e + i] = (double)tmp[i];
        e += n;
      }
    });
  for (auto &th : pool) th.join();
  ::close(fd);
  for (int b : bad) if (b) die("short read from " + path);
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

// Double-buffered source of host snapshots: a reader thread fills page-locked buffers ahead of the consumer.
struct Feed {
  static constexpr int NB = 2;
  double *buf[NB] = {nullptr, nullptr};
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  long produced = 0, consumed = 0, total = 0;
  bool pinned = false, failed = false;
  std::string error;
  ~Feed() {
    { std::lock_guard<std::mutex> lk(m); consumed = total; }
    cv.notify_all();
    if (th.joinable()) th.join();
    for (double *b : buf) { if (pinned) ftkb_host_free(b); else std::free(b); }
  }
  void start(long T, size_t bytes, std::function<void(long, double *)> fill) {
    total = T;
    pinned = true;
    for (int i = 0; i < NB; i++) {
      void *p = nullptr;
      if (ftkb_host_alloc(bytes, &p) != FTKB_OK) { pinned = false; break; }
      buf[i] = static_cast<double *>(p);
    }
    if (!pinned)      // no device / no pinned memory: plain buffers (the tracker will fail loudly later if there is no device)
      for (int i = 0; i < NB; i++) { if (buf[i]) ftkb_host_free(buf[i]); buf[i] = static_cast<double *>(std::malloc(bytes)); }
    th = std::thread([this, fill] {
      for (long k = 0; k < total; k++) {
        {
          std::unique_lock<std::mutex> lk(m);
          cv.wait(lk, [&] { return k - consumed < NB || consumed >= total; });
          if (consumed >= total) return;
        }
        try { fill(k, buf[k % NB]); }
        catch (const std::exception &e) { std::lock_guard<std::mutex> lk(m); failed = true; error = e.what(); }
        { std::lock_guard<std::mutex> lk(m); produced = k + 1; }
        cv.notify_all();
      }
    });
  }
  double *get(long k) {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return produced > k || failed; });
    if (failed) throw std::runtime_error(error);
    return buf[k % NB];
  }
  void release(long k) {
    { std::lock_guard<std::mutex> lk(m); consumed = k + 1; }
    cv.notify_all();
  }
};

int main(int argc, char **argv) {
  Options o = parse(argc, argv);
  if (o.help || argc == 1) { usage(); return 0; }
  if (o.feature != "cp" && o.feature != "critical_point") die("only '-f cp' (critical points on regular grids) is implemented");
  if (o.accelerator != "cuda") die("invalid '--accelerator': this build runs on CUDA (sm_100a) only, there is no CPU path");
  if (o.output.empty()) die("Missing '--output'.");
  if (o.synthetic.empty() == o.input.empty()) die("give exactly one of --synthetic and --input");

  // stream description (stream.hh:258-440 defaults)
  int nd = 2, nv = 1;
  long dims[3] = {32, 32, 32}, T = 32;
  int syn = -1;   // FTKB_SYN_* for --device-generators
  if (!o.synthetic.empty()) {
    const std::string &s = o.synthetic;
    if (s == "woven") { syn = FTKB_SYN_WOVEN; }
    else if (s == "moving_extremum_2d") { dims[0] = dims[1] = 21; syn = FTKB_SYN_MOVING_EXTREMUM; if (o.x0.empty()) o.x0 = {10.0, 10.0}; if (o.dir.empty()) o.dir = {0.1, 0.1}; }
    else if (s == "moving_extremum_3d") { nd = 3; dims[0] = dims[1] = dims[2] = 21; syn = FTKB_SYN_MOVING_EXTREMUM; if (o.x0.empty()) o.x0 = {10, 10, 10}; if (o.dir.empty()) o.dir = {0.1, 0.11, 0.1}; }
    else if (s == "double_gyre") { dims[0] = 64; dims[1] = 32; T = 50; nv = 2; syn = FTKB_SYN_DOUBLE_GYRE; }
    else if (s == "merger_2d") { T = 100; syn = FTKB_SYN_MERGER; }
    else if (s == "tornado") { nd = 3; nv = 3; syn = FTKB_SYN_TORNADO; }                     // stream.hh:431,1560-1567: 32^3 x 32
    else die("synthetic case not available: " + s);
    if (syn == FTKB_SYN_MOVING_EXTREMUM && ((int)o.x0.size() != nd || (int)o.dir.size() != nd)) die("invalid x0 / dir");
  } else {
    if (o.width < 0 || o.height < 0 || o.timesteps < 0) die("--input needs --width, --height[, --depth] and --timesteps");
    nd = o.depth > 0 ? 3 : 2;
    if (!o.var.empty()) nv = (int)parse_list(o.var).size();
    if (nv != 1 && nv != nd) die("--var must name 1 variable (scalar field) or as many as spatial dimensions (vector field)");
    if (o.input_format != "float32" && o.input_format != "float64") die("--input-format must be float32 or float64");
  }
  if (o.width > 0) dims[0] = o.width;
  if (o.height > 0) dims[1] = o.height;
  if (o.depth > 0) { if (nd != 3) die("--depth is valid only for 3D data"); dims[2] = o.depth; }
  if (o.timesteps > 0) T = o.timesteps;

  const double t_start = now();
  std::unique_ptr<critical_point_tracker_regular> tr;
  if (nd == 2) tr.reset(new critical_point_tracker_2d_regular()); else tr.reset(new critical_point_tracker_3d_regular());
  // json_interface.hh:634-656: scalar -> derived gradient + symmetric Jacobian on [2, D-2]; vector -> given, non-symmetric, [1, D-2]
  std::vector<int> lo(nd), sz(nd), zero(nd, 0), full(nd);
  for (int i = 0; i < nd; i++) { lo[i] = nv == 1 ? 2 : 1; sz[i] = (int)dims[i] - (nv == 1 ? 3 : 2); full[i] = (int)dims[i]; }
  try {
    tr->set_domain(lattice(lo, sz));
    tr->set_array_domain(lattice(zero, full));
    if (nv == 1) {
      tr->set_scalar_field_source(SOURCE_GIVEN); tr->set_vector_field_source(SOURCE_DERIVED); tr->set_jacobian_field_source(SOURCE_DERIVED);
      tr->set_jacobian_symmetric(true);
    } else {
      tr->set_scalar_field_source(SOURCE_NONE); tr->set_vector_field_source(SOURCE_GIVEN); tr->set_jacobian_field_source(SOURCE_DERIVED);
      tr->set_jacobian_symmetric(false);
      tr->set_scalar_components({});
    }
    tr->set_number_of_threads(o.nthreads);
    tr->use_accelerator(o.accelerator);
    tr->set_device_ids(o.devices.empty() ? std::vector<int>{o.device} : o.devices);
    tr->set_time_chunk(o.time_chunk);
    if (!o.type_filter.empty()) tr->set_type_filter(parse_type_filter(o.type_filter));
    if (o.compute_degrees) tr->set_enable_computing_degrees(true);
    if (o.stream) tr->set_enable_streaming_trajectories(true);      // json_interface.hh:324-325
    if (o.no_robust) tr->set_enable_robust_detection(false);
    tr->initialize();
    const double t_init = now();

    size_t nvert = 1;
    for (int i = 0; i < nd; i++) nvert *= (size_t)dims[i];
    std::vector<size_t> shape;
    if (nv > 1) shape.push_back(nv);
    for (int i = 0; i < nd; i++) shape.push_back((size_t)dims[i]);
    // host-side sources (files, host generators): two page-locked buffers and a reader thread -- snapshot k+1 is read (or
    // generated) and converted while snapshot k travels to the device and is swept (the reference overlaps I/O and compute the
    // same way, ndarray/stream.hh:1607-1699)
    Feed feed;
    // raw float32 series on one device: the bytes go to the device as they are (half the PCIe traffic) and are widened there
    const bool raw_f32 = syn < 0 && o.input_format == "float32" && o.devices.size() <= 1;
    double t_wait = 0, t_push = 0, t_step = 0;
    if (!(syn >= 0 && o.device_generators))
      feed.start(T, nvert * nv * (raw_f32 ? sizeof(float) : sizeof(double)), [&](long k, double *dst) {
        if (syn == FTKB_SYN_WOVEN) gen_woven(dims[0], dims[1], T == 1 ? 0.0 : double(k) / (T - 1), dst);        // stream.hh:1468-1480
        else if (syn == FTKB_SYN_MERGER) gen_merger(dims[0], dims[1], double(k) * 0.1, dst);                   // stream.hh:1540
        else if (syn == FTKB_SYN_DOUBLE_GYRE) gen_double_gyre(dims[0], dims[1], k * o.time_scale, dst);        // stream.hh:1542-1555
        else if (syn == FTKB_SYN_TORNADO) gen_tornado(dims[0], dims[1], dims[2], (int)k, dst);                  // stream.hh:1560-1567
        else if (syn == FTKB_SYN_MOVING_EXTREMUM) gen_moving_extremum(nd, dims, o.x0.data(), o.dir.data(), double(k), dst);
        else read_raw(o, k, nvert * nv, o.input_format == "float32", dst, raw_f32);
      });
    const double t_feed_ready = now();              // page-locked buffers allocated, reader thread running
    for (long k = 0; k < T; k++) {
      if (o.verbose) std::fprintf(stderr, "current_timestep=%ld\n", k);
      if (syn >= 0 && o.device_generators) {
        std::vector<double> p;
        double t = double(k);
        if (syn == FTKB_SYN_WOVEN) t = T == 1 ? 0.0 : double(k) / (T - 1);
        else if (syn == FTKB_SYN_MERGER) t = double(k) * 0.1;
        else if (syn == FTKB_SYN_DOUBLE_GYRE) { t = k * o.time_scale; p = {0.1, M_PI * 2, 0.25}; }
        else { p = o.x0; p.insert(p.end(), o.dir.begin(), o.dir.end()); }
        tr->push_synthetic_snapshot(syn, p, t);
      } else {
        const double t0 = now();
        double *buf = feed.get(k);                   // filled by the reader thread while the previous snapshot was pushed and swept
        const double t1 = now();
        t_wait += t1 - t0;
        if (raw_f32) {
          const ndarray<float> a = ndarray<float>::wrap(reinterpret_cast<const float *>(buf), shape);
          if (nv == 1) tr->push_scalar_field_snapshot(a); else tr->push_vector_field_snapshot(a);
        } else {
          const ndarray<double> a = ndarray<double>::wrap(buf, shape);
          if (nv == 1) tr->push_scalar_field_snapshot(a); else tr->push_vector_field_snapshot(a);
        }
        feed.release(k);                             // (the push copies before it returns)
        t_push += now() - t1;
      }
      const double t2 = now();
      if (k != 0) tr->advance_timestep();          // json_interface.hh:699-706
      if (k == T - 1) tr->update_timestep();
      t_step += now() - t2;
    }
    const double t_compute = now();
    const double t_feed_alloc = t_feed_ready - t_init;
    if (o.output_type == "traced" || o.output_type == "sliced") {
      tr->finalize();
      if (!o.post_process.empty()) tr->post_process(o.post_process);     // feature_curve_set_post_processor_t(ops).filter(trajs)
    }
    const double t_final = now();

    std::string fmt = o.output_format;
    if (fmt == "auto") fmt = (o.output.size() > 5 && o.output.substr(o.output.size() - 5) == ".json") ? "json" : "text";
    if (o.output_type == "traced") {
      if (fmt == "json") tr->write_traced_critical_points_json(o.output); else if (fmt == "text") tr->write_traced_critical_points_text(o.output);
      else if (fmt == "binary") tr->write_traced_critical_points_binary(o.output); else die("unsupported --output-format " + fmt);
    } else if (o.output_type == "discrete") {
      if (fmt == "json") tr->write_critical_points_json(o.output); else if (fmt == "text") tr->write_critical_points_text(o.output);
      else if (fmt == "binary") tr->write_critical_points_binary(o.output); else die("unsupported --output-format " + fmt);
    } else if (o.output_type == "sliced") {       // json_interface.hh:805-811,584-592: one file per timestep, series_filename(pattern, k)
      if (fmt != "text") die("sliced output is written as text");
      if (o.output.find('%') == std::string::npos) die("--output-type sliced needs a printf pattern in --output (e.g. sliced-%03d.txt)");
      tr->slice_traced_critical_points();
      for (const auto &kv : tr->get_sliced_critical_points()) {
        char name[4096];
        std::snprintf(name, sizeof(name), o.output.c_str(), kv.first);
        tr->write_sliced_critical_points_text(kv.first, std::string(name));
      }
    } else die("unsupported --output-type " + o.output_type + " (traced | discrete | sliced)");

    if (o.timing) {   // same line as json_interface.hh:718-723, plus the device-side split
      const ftkb_stats st = tr->stats();
      std::fprintf(stderr, "t_init=%f, t_compute=%f, t_finalize=%f\n", t_init - t_start, t_compute - t_init, t_final - t_compute);
      std::fprintf(stderr, "t_input_buffers=%f, t_wait_input=%f, t_push=%f, t_step=%f (inside t_compute; %ld snapshots)\n", t_feed_alloc, t_wait, t_push, t_step, T);
      std::fprintf(stderr, "simplices_tested=%llu, punctured=%llu, kernel_launches=%llu, ms_scan=%f, ms_test=%f, ms_derive=%f, h2d_bytes=%llu\n",
                   (unsigned long long)st.simplices_tested, (unsigned long long)st.points, (unsigned long long)st.kernel_launches, st.ms_scan, st.ms_test, st.ms_derive,
                   (unsigned long long)st.h2d_bytes);
    }
  } catch (const std::exception &e) {
    die(e.what());
  }
  return 0;
}
