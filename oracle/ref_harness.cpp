// TEST INFRASTRUCTURE ONLY -- not part of the shipped product.
//
// Driver for the UNMODIFIED reference (hguo/ftk) CPU critical-point tracker.  It is compiled
// against the headers where they lie under /root/reference/include (see oracle/build_ref.sh)
// into oracle/_ref/ftk_ref_oracle.  Nothing under /root/reference is copied: this file only
// *calls* the reference API the way its own front ends do
//   (python/pyftk.cpp:93-117, include/ftk/filters/json_interface.hh:606-706).
//
// Uses: (1) generating tests/golden/*.ftkg fixtures (tests/golden/make_golden.py),
//       (2) validating the C restatement oracle/cp_oracle.c on arbitrary inputs,
//       (3) the CPU baseline of bench.py (`--impl reference`, cpu_baseline.kind == "reference").
//
// Usage:
//   ftk_ref_oracle --nd 2|3 --nv 1|2|3 --dims W H [D] --nt T
//                  (--gen NAME [--p a b c ...] | --input raw.f64)
//                  [--domain lb0 ub0 lb1 ub1 [lb2 ub2]] [--symmetric 0|1] [--nthreads N]
//                  [--no-trace] [--out file.ftkg] [--dump-input raw.f64] [--quiet]
//                  [--binary-discrete file] [--binary-traced file]   the reference's own binary archives (DIY serialization)
//                  [--out-stream file]   trajectories of a second tracker run with streaming trajectories
//                                        (trace_critical_points_online per step), as indices into the punctured simplices
//                  [--post OPS --out-curves file.ftkc]   trajectory post-processing by the reference's own curve code:
//                  OPS as feature_curve_set_post_processor_t takes them, plus legacy[:thr[:discard[:velocity]]] = the
//                  call sequence of json_interface::post_process() (json_interface.hh:758-800) on the same methods
//   raw.f64 holds T consecutive snapshots, each (nv, W, H[, D]) float64 with dim 0 fastest.

#include <ftk/filters/critical_point_tracker_2d_regular.hh>
#include <ftk/filters/critical_point_tracker_3d_regular.hh>
#include <ftk/ndarray/synthetic.hh>
#include <ftk/filters/feature_curve_set_post_processor.hh>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using ftk::ndarray;
typedef ftk::simplicial_regular_mesh_element element_t;

struct args_t {
  int nd = 2, nv = 1, W = 0, H = 0, D = 1, T = 0, nthreads = 0, start_timestep = 0;
  int symmetric = -1;
  bool trace = true, quiet = false, have_domain = false;
  int dom[6] = {0, 0, 0, 0, 0, 0};
  std::string gen, input, out, dump_input, post, out_curves, bin_traced, bin_discrete, out_stream, coords, coords_file;
  std::vector<double> p;
};

static ndarray<double> make_snapshot(const args_t &a, int k, FILE *fin)
{
  std::vector<size_t> shape;
  if (a.nv > 1) shape.push_back(a.nv);
  shape.push_back(a.W); shape.push_back(a.H);
  if (a.nd == 3) shape.push_back(a.D);

  if (fin) {
    ndarray<double> arr(shape);
    if (fread(arr.data(), sizeof(double), arr.nelem(), fin) != arr.nelem()) {
      fprintf(stderr, "short read on input, k=%d\n", k); exit(2);
    }
    if (a.nv > 1) arr.set_multicomponents();
    return arr;
  }
  const std::string &g = a.gen;
  auto P = [&](size_t i, double dflt) { return i < a.p.size() ? a.p[i] : dflt; };
  if (g == "woven") {            // pyftk synthesizers.spiral_woven slice k
    const double t = (double(k) / (a.T - 1)) + 1e-4;
    return ftk::synthetic_woven_2D<double>(a.W, a.H, t);
  } else if (g == "woven_cli") { // `ftk --synthetic woven` (stream.hh request_timestep_synthetic_woven)
    const double t = a.T == 1 ? 0.0 : double(k) / (a.T - 1);
    return ftk::synthetic_woven_2D<double>(a.W, a.H, t);
  } else if (g == "merger") {
    return ftk::synthetic_merger_2D<double>(a.W, a.H, double(k) * 0.1);
  } else if (g == "moving_extremum" && a.nd == 2) {
    const double x0[2] = {P(0, 10), P(1, 10)}, dir[2] = {P(2, 0.1), P(3, 0.1)};
    return ftk::synthetic_moving_extremum<double, 2>({(size_t)a.W, (size_t)a.H}, x0, dir, double(k));
  } else if (g == "moving_extremum" && a.nd == 3) {
    const double x0[3] = {P(0, 10), P(1, 10), P(2, 10)}, dir[3] = {P(3, 0.1), P(4, 0.11), P(5, 0.1)};
    return ftk::synthetic_moving_extremum<double, 3>({(size_t)a.W, (size_t)a.H, (size_t)a.D}, x0, dir, double(k));
  } else if (g == "double_gyre") {
    return ftk::synthetic_double_gyre<double>(a.W, a.H, k * P(0, 0.1), false, 0.1, M_PI * 2, 0.25);
  } else if (g == "abc") {       // unsteady ABC: A(k) = sqrt(3) + 0.5 (k/64) sin(pi k/64)  
    const double A = std::sqrt(3.0) + 0.5 * (double(k) / 64.0) * std::sin(M_PI * double(k) / 64.0);
    return ftk::synthetic_abc_flow<double>(a.W, a.H, a.D, A, std::sqrt(2.0), 1.0);
  } else if (g == "tornado") {
    return ftk::synthetic_tornado<double>(a.W, a.H, a.D, k);
  }
  fprintf(stderr, "unknown generator %s\n", g.c_str()); exit(2);
}

template <typename tracker_t>
static int run(const args_t &a)
{
  diy::mpi::communicator comm;
  tracker_t tr(comm);
  const int nd = a.nd;
  const size_t dims[3] = {(size_t)a.W, (size_t)a.H, (size_t)a.D};

  if (a.nthreads > 0) tr.set_number_of_threads(a.nthreads);

  std::vector<size_t> dst, dsz, ast, asz;
  const bool scalar = a.nv == 1;
  for (int i = 0; i < nd; i ++) {
    size_t lb = scalar ? 2 : 1, ub = dims[i] - 2; // json_interface.hh:639-654: lattice({2..},{D-3..}) / ({1..},{D-2..}) => inclusive ub = D-2
    if (a.have_domain) { lb = a.dom[2*i]; ub = a.dom[2*i+1]; }
    dst.push_back(lb); dsz.push_back(ub - lb + 1);
    ast.push_back(0);  asz.push_back(dims[i]);
  }
  if (scalar) {
    tr.set_scalar_field_source(ftk::SOURCE_GIVEN);
    tr.set_vector_field_source(ftk::SOURCE_DERIVED);
    tr.set_jacobian_field_source(ftk::SOURCE_DERIVED);
    tr.set_jacobian_symmetric(a.symmetric < 0 ? true : a.symmetric != 0);
  } else {
    tr.set_scalar_field_source(ftk::SOURCE_NONE);
    tr.set_vector_field_source(ftk::SOURCE_GIVEN);
    tr.set_jacobian_field_source(ftk::SOURCE_DERIVED);
    tr.set_jacobian_symmetric(a.symmetric < 0 ? false : a.symmetric != 0);
  }
  tr.set_domain(ftk::lattice(dst, dsz));
  tr.set_array_domain(ftk::lattice(ast, asz));
  if (!a.coords.empty()) {   // regular_tracker.hh:38-40
    FILE *fc = fopen(a.coords_file.c_str(), "rb"); if (!fc) { perror("coords-file"); return 2; }
    std::vector<double> cd;
    double tmp;
    while (fread(&tmp, 8, 1, fc) == 1) cd.push_back(tmp);
    fclose(fc);
    if (a.coords == "bounds") tr.set_coords_bounds(cd);
    else if (a.coords == "rectilinear") {
      std::vector<ftk::ndarray<double>> rc;
      size_t off = 0;
      for (int i = 0; i < nd; i ++) { ftk::ndarray<double> r; r.reshape({dims[i]}); for (size_t k = 0; k < dims[i]; k ++) r[k] = cd[off + k]; off += dims[i]; rc.push_back(r); }
      tr.set_coords_rectilinear(rc);
    } else if (a.coords == "explicit") {
      const size_t wh = dims[0] * dims[1], nc = cd.size() / wh;
      ftk::ndarray<double> e; e.reshape({nc, dims[0], dims[1]});
      for (size_t k = 0; k < nc * wh; k ++) e[k] = cd[k];
      tr.set_coords_explicit(e);
    } else { fprintf(stderr, "unknown --coords %s\n", a.coords.c_str()); return 2; }
  }
  tr.initialize();
  if (a.start_timestep) tr.set_current_timestep(a.start_timestep);   // tracker.hh:40 (a time slab, or a run resumed at t > 0)

  tracker_t trs(comm);           // streaming twin: same configuration, trajectories grown online after every step
  const bool streaming = !a.out_stream.empty();
  if (streaming) {
    if (a.nthreads > 0) trs.set_number_of_threads(a.nthreads);
    trs.set_enable_streaming_trajectories(true);
    if (scalar) {
      trs.set_scalar_field_source(ftk::SOURCE_GIVEN); trs.set_vector_field_source(ftk::SOURCE_DERIVED); trs.set_jacobian_field_source(ftk::SOURCE_DERIVED);
      trs.set_jacobian_symmetric(a.symmetric < 0 ? true : a.symmetric != 0);
    } else {
      trs.set_scalar_field_source(ftk::SOURCE_NONE); trs.set_vector_field_source(ftk::SOURCE_GIVEN); trs.set_jacobian_field_source(ftk::SOURCE_DERIVED);
      trs.set_jacobian_symmetric(a.symmetric < 0 ? false : a.symmetric != 0);
    }
    trs.set_domain(ftk::lattice(dst, dsz));
    trs.set_array_domain(ftk::lattice(ast, asz));
    trs.initialize();
  }

  FILE *fin = NULL, *fdump = NULL;
  if (!a.input.empty()) { fin = fopen(a.input.c_str(), "rb"); if (!fin) { perror("input"); return 2; } }
  if (!a.dump_input.empty()) { fdump = fopen(a.dump_input.c_str(), "wb"); if (!fdump) { perror("dump"); return 2; } }

  typedef std::chrono::high_resolution_clock clk;
  double t_gen = 0, t_push = 0, t_sweep = 0, t_final = 0;
  for (int k = 0; k < a.T; k ++) {
    auto c0 = clk::now();
    ndarray<double> s = make_snapshot(a, k, fin);
    if (fdump) fwrite(s.data(), sizeof(double), s.nelem(), fdump);
    auto c1 = clk::now();
    if (scalar) tr.push_scalar_field_snapshot(s);
    else tr.push_vector_field_snapshot(s);
    auto c2 = clk::now();
    if (k != 0) tr.advance_timestep();
    if (k == a.T - 1) tr.update_timestep();
    auto c3 = clk::now();
    if (streaming) {
      if (scalar) trs.push_scalar_field_snapshot(s); else trs.push_vector_field_snapshot(s);
      if (k != 0) trs.advance_timestep();
      if (k == a.T - 1) trs.update_timestep();
    }
    t_gen += std::chrono::duration<double>(c1 - c0).count();
    t_push += std::chrono::duration<double>(c2 - c1).count();
    t_sweep += std::chrono::duration<double>(c3 - c2).count();
  }
  if (fin) fclose(fin);
  if (fdump) fclose(fdump);

  if (!a.bin_discrete.empty()) tr.write_critical_points_binary(a.bin_discrete);
  // discrete critical points (std::map order == element operator<, x-first lexicographic)
  const auto pts = tr.get_discrete_critical_points(); // copy: finalize() may clear on non-root
  std::map<unsigned long long, size_t> tag2idx;
  { size_t i = 0; for (const auto &kv : pts) tag2idx[kv.second.tag] = i ++; }
  const bool tags_unique = tag2idx.size() == pts.size();

  size_t ntraj = 0;
  std::vector<std::vector<size_t>> trajs; std::vector<int> loops;
  if (a.trace) {
    auto c0 = clk::now();
    tr.finalize();
    t_final = std::chrono::duration<double>(clk::now() - c0).count();
    if (!tags_unique) { fprintf(stderr, "tags not unique; cannot index trajectories\n"); return 3; }
    if (!a.bin_traced.empty()) tr.write_traced_critical_points_binary(a.bin_traced);
    for (const auto &kv : tr.get_traced_critical_points()) {
      std::vector<size_t> idx;
      for (const auto &cp : kv.second) idx.push_back(tag2idx.at(cp.tag));
      trajs.push_back(idx); loops.push_back(kv.second.loop ? 1 : 0);
    }
    ntraj = trajs.size();
  }

  if (a.trace && !a.out_curves.empty()) {
    auto &set = tr.get_traced_critical_points();
    for (const std::string &op : ftk::split(a.post, ",")) {
      if (op.empty()) continue;
      if (op.rfind("legacy", 0) == 0) {          // json_interface::post_process(), same calls in the same order
        const auto f = ftk::split(op, ":");
        const double thr = f.size() > 1 ? atof(f[1].c_str()) : 0.0;
        const bool disc = f.size() > 2 && atoi(f[2].c_str()), vel = f.size() > 3 && atoi(f[3].c_str());
        set.foreach([](ftk::feature_curve_t &t) { t.smooth_ordinal_types(); t.smooth_interval_types(); t.rotate(); t.update_statistics(); });
        if (thr > 0) set.filter([&](const ftk::feature_curve_t &t) { return !(t.tmax - t.tmin < thr); });
        set.split_all();
        if (disc) set.foreach([](ftk::feature_curve_t &t) { t.discard_interval_points(); });
        set.foreach([](ftk::feature_curve_t &t) { t.reorder(); t.adjust_time(); t.update_statistics(); });
        if (vel) set.foreach([](ftk::feature_curve_t &t) { t.discard_interval_points(); t.derive_velocity(); t.update_statistics(); });
      } else if (op.rfind("duration_pruning:", 0) == 0) {
        const double thr = atof(op.c_str() + 17);
        if (thr > 0) set.filter([&](const ftk::feature_curve_t &t) { return !(t.tmax - t.tmin < thr); });
      } else if (op.rfind("intercept:", 0) == 0) {                    // feature_curve_set_t::intercept(t0, t1)
        const auto f = ftk::split(op, ":");
        const ftk::feature_curve_set_t cut = set.intercept(atoi(f[1].c_str()), atoi(f[2].c_str()));
        set = cut;
      } else if (op == "discard_degenerate_points") {
        set.foreach([](ftk::feature_curve_t &t) { t.discard_degenerate_points(); t.update_statistics(); });
      } else if (op == "update_statistics") {
        set.foreach([](ftk::feature_curve_t &t) { t.update_statistics(); });
      } else {
        ftk::feature_curve_set_post_processor_t pp(op);
        pp.filter(set);
      }
    }
    FILE *fc = fopen(a.out_curves.c_str(), "wb"); if (!fc) { perror("out-curves"); return 2; }
    const uint32_t magic = 0x434b5446 /*FTKC*/, version = 1;
    const uint64_t nc = set.size();
    fwrite(&magic, 4, 1, fc); fwrite(&version, 4, 1, fc); fwrite(&nc, 8, 1, fc);
    for (const auto &kv : set) {
      const auto &t = kv.second;
      const int32_t h[4] = {kv.first, t.loop ? 1 : 0, t.complete ? 1 : 0, (int32_t)t.consistent_type};
      const uint64_t n = t.size();
      const double st[13] = {t.tmin, t.tmax, t.bbmin[0], t.bbmin[1], t.bbmin[2], t.bbmax[0], t.bbmax[1], t.bbmax[2],
                             t.min[0], t.max[0], t.persistence[0], t.vmmin, t.vmmax};
      fwrite(h, 4, 4, fc); fwrite(&n, 8, 1, fc); fwrite(st, 8, 13, fc);
      for (const auto &cp : t) {
        const uint64_t idx = tag2idx.at(cp.tag);
        const int32_t q[4] = {(int32_t)cp.type, cp.ordinal ? 1 : 0, (int32_t)cp.timestep, (int32_t)cp.id};
        const double d[8] = {cp.x[0], cp.x[1], cp.x[2], cp.t, cp.scalar[0], cp.v[0], cp.v[1], cp.v[2]};
        fwrite(&idx, 8, 1, fc); fwrite(q, 4, 4, fc); fwrite(d, 8, 8, fc);
      }
    }
    // sliced critical points (critical_point_tracker.hh:819-835): ordinal points of the traced curves per timestep
    tr.slice_traced_critical_points();
    const auto &sliced = tr.get_sliced_critical_points();
    const uint64_t ns = sliced.size();
    fwrite(&ns, 8, 1, fc);
    for (const auto &kv : sliced) {
      const int32_t h[2] = {kv.first, 0};
      const uint64_t n = kv.second.size();
      fwrite(h, 4, 2, fc); fwrite(&n, 8, 1, fc);
      for (const auto &cp : kv.second) { const uint64_t idx = tag2idx.at(cp.tag); fwrite(&idx, 8, 1, fc); }
    }
    fclose(fc);
  }

  if (streaming) {
    trs.finalize();
    FILE *fs = fopen(a.out_stream.c_str(), "wb"); if (!fs) { perror("out-stream"); return 2; }
    const auto &set = trs.get_traced_critical_points();
    const uint64_t nt = set.size();
    fwrite(&nt, 8, 1, fs);
    for (const auto &kv : set) {
      const uint64_t h[4] = {(uint64_t)kv.first, kv.second.size(), kv.second.loop ? 1u : 0u, kv.second.complete ? 1u : 0u};
      fwrite(h, 8, 4, fs);
      for (const auto &cp : kv.second) { const uint64_t idx = tag2idx.at(cp.tag); fwrite(&idx, 8, 1, fs); }
    }
    fclose(fs);
  }

  // simplices enumerated: N_core * (n_ord * T + n_int * (T-1))   
  double ncore = 1; for (int i = 0; i < nd; i ++) ncore *= double(dsz[i]);
  const int n_ord = nd == 2 ? 2 : 6, n_int = nd == 2 ? 10 : 54;
  const double nsimplices = ncore * (double(n_ord) * a.T + double(n_int) * (a.T - 1));

  printf("{\"npoints\": %zu, \"ntraj\": %zu, \"simplices\": %.0f, \"t_gen\": %.6f, \"t_push\": %.6f, "
         "\"t_sweep\": %.6f, \"t_finalize\": %.6f, \"nthreads\": %d}\n",
         pts.size(), ntraj, nsimplices, t_gen, t_push, t_sweep, t_final, tr.get_number_of_threads());

  if (!a.out.empty()) {
    FILE *fo = fopen(a.out.c_str(), "wb"); if (!fo) { perror("out"); return 2; }
    const uint32_t magic = 0x474b5446 /*FTKG*/, version = 1, und = nd;
    const uint64_t np = pts.size(), nt = ntraj;
    fwrite(&magic, 4, 1, fo); fwrite(&version, 4, 1, fo); fwrite(&und, 4, 1, fo);
    const uint32_t traced = a.trace ? 1 : 0; fwrite(&traced, 4, 1, fo);
    fwrite(&np, 8, 1, fo); fwrite(&nt, 8, 1, fo);
    for (const auto &kv : pts) {
      int32_t rec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < nd + 1; i ++) rec[i] = kv.first.corner[i];  // x, y, [z,] t
      if (nd == 2) { rec[3] = rec[2]; rec[2] = 0; }                       // always (x, y, z, t)
      rec[4] = kv.first.type; rec[5] = kv.second.ordinal ? 1 : 0; rec[6] = kv.second.timestep;
      rec[7] = (int32_t)kv.second.type;
      fwrite(rec, 4, 8, fo);
      const double d[5] = {kv.second.x[0], kv.second.x[1], kv.second.x[2], kv.second.t, kv.second.scalar[0]};
      fwrite(d, 8, 5, fo);
    }
    for (size_t i = 0; i < ntraj; i ++) {
      const uint64_t len = trajs[i].size(), loop = loops[i];
      fwrite(&len, 8, 1, fo); fwrite(&loop, 8, 1, fo);
      for (auto j : trajs[i]) { const uint64_t v = j; fwrite(&v, 8, 1, fo); }
    }
    fclose(fo);
  }
  return 0;
}

int main(int argc, char **argv)
{
  diy::mpi::environment env;
  args_t a;
  for (int i = 1; i < argc; i ++) {
    const std::string s = argv[i];
    auto next = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", s.c_str()); exit(2); } return argv[++i]; };
    if (s == "--nd") a.nd = atoi(next());
    else if (s == "--nv") a.nv = atoi(next());
    else if (s == "--dims") { a.W = atoi(next()); a.H = atoi(next()); if (a.nd == 3) a.D = atoi(next()); }
    else if (s == "--nt") a.T = atoi(next());
    else if (s == "--start-timestep") a.start_timestep = atoi(next());
    else if (s == "--gen") a.gen = next();
    else if (s == "--input") a.input = next();
    else if (s == "--out") a.out = next();
    else if (s == "--dump-input") a.dump_input = next();
    else if (s == "--symmetric") a.symmetric = atoi(next());
    else if (s == "--nthreads") a.nthreads = atoi(next());
    else if (s == "--no-trace") a.trace = false;
    else if (s == "--binary-traced") a.bin_traced = next();      // tracker.write_traced_critical_points_binary (DIY archive)
    else if (s == "--binary-discrete") a.bin_discrete = next();  // tracker.write_critical_points_binary
    else if (s == "--out-stream") a.out_stream = next();         // a second tracker with set_enable_streaming_trajectories(true)
    else if (s == "--coords") a.coords = next();
    else if (s == "--coords-file") a.coords_file = next();
    else if (s == "--post") a.post = next();
    else if (s == "--out-curves") a.out_curves = next();
    else if (s == "--quiet") a.quiet = true;
    else if (s == "--domain") { a.have_domain = true; for (int j = 0; j < 2 * a.nd; j ++) a.dom[j] = atoi(next()); }
    else if (s == "--p") { while (i + 1 < argc && strncmp(argv[i+1], "--", 2) != 0) a.p.push_back(atof(argv[++i])); }
    else { fprintf(stderr, "unknown argument %s\n", s.c_str()); return 2; }
  }
  if (a.W <= 0 || a.H <= 0 || a.T <= 0 || (a.gen.empty() && a.input.empty())) {
    fprintf(stderr, "usage: see header of oracle/ref_harness.cpp\n"); return 2;
  }
  if (a.quiet) { if (!freopen("/dev/null", "w", stderr)) return 2; }
  if (a.nd == 2) return run<ftk::critical_point_tracker_2d_regular>(a);
  else return run<ftk::critical_point_tracker_3d_regular>(a);
}
