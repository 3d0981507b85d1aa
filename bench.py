#!/usr/bin/env python
"""bench.py -- space-time simplices tested per second (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c5] [--impl reference]

One "step" = one advance_timestep of the tracker on one new snapshot: field derivation (gradient +
min non-zero |v|), quantisation factor, ordinal sweep at t and interval sweep over [t, t+1]
(scan + per-simplex test kernels), i.e. 12 (2D) / 60 (3D) enumerated simplices per domain corner.

Default workload = BASELINE.json configs[1]: moving-extremum 2D scalar field 8192 x 8192, 64 distinct
timesteps resident in HBM (the sweep ping-pongs through them when K > 63).  `value` is measured
with the snapshots already resident in HBM (borrowed in place through the C ABI); `e2e` repeats the
measurement through the same C-ABI calls with pinned HOST buffers (H2D of every snapshot inside
the timed region).  N > 1: one process per GPU, each rank owns a contiguous time slab of K steps
(weak scaling), receives its one-layer halo from the next rank with NCCL send/recv inside the
timed region, and verifies the running quantisation factor against the all-gathered per-layer
resolutions.  Finalize (union-find + trace ordering) is reported separately.

--impl reference times the unmodified reference CPU tracker (oracle/_ref/ftk_ref_oracle, built from
/root/reference by oracle/build_ref.sh) on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (workload label, dims, generator params x0+dir, field, distinct timesteps)
    "c2": ("moving_extremum_2d_scalar_8192x8192x64", [8192, 8192], [4096.3, 4095.7, 0.1, 0.1], 64),
    "c3": ("moving_extremum_3d_scalar_512x512x512x32", [512, 512, 512], [256.3, 255.7, 256.1, 0.1, 0.11, 0.1], 32),
    "c2s": ("moving_extremum_2d_scalar_2048x2048x16", [2048, 2048], [1024.3, 1023.7, 0.1, 0.1], 16),
    # vector input (field GIVEN, Jacobian derived at punctured simplices): BASELINE configs[3] and configs[4]
    "c4": ("unsteady_abc_3d_vector_1024x1024x1024x64", [1024, 1024, 1024], None, 4),      # 25.8 GB per layer: 4 distinct layers resident
    "c5": ("double_gyre_2d_vector_16384x8192x256", [16384, 8192], None, 32),
    "c4s": ("unsteady_abc_3d_vector_256x256x256x64", [256, 256, 256], None, 8),
}
VECTOR_CONFIGS = ("c4", "c5", "c4s")
METRIC = "space_time_simplices_tested_per_sec"
UNIT = "simplices/s"


def tri(g, period):
    """ping-pong index 0..period-1..0 so that the extremum keeps moving continuously"""
    m = 2 * (period - 1)
    g = g % m
    return g if g < period else m - g


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ---------------------------------------------------------------------------------------------------
def reference_arm(args, rank):
    """the unmodified reference CPU tracker on a bounded sample of the workload (rank 0 only)"""
    if rank != 0:
        return
    label, dims, params, _ = CONFIGS[args.config]
    nd = len(dims)
    binary = os.path.join(ROOT, "oracle", "_ref", "ftk_ref_oracle")
    kind = "reference"
    cores = os.cpu_count() or 1
    # one step = the reference's tracker loop over 2 snapshots of a reduced extent of the same generator
    sdims = [512, 512] if nd == 2 else [48, 48, 48]
    sp = [d / 2 + 0.3 for d in sdims] + list(params[nd:])
    ncore = 1
    for d in sdims:
        ncore *= d - 3
    simplices = ncore * ((12 + 2) if nd == 2 else (60 + 6))   # advance (ordinal + interval) + final ordinal sweep
    sample = f"{'x'.join(map(str, sdims))}x2 timesteps of the same generator per step, {cores} threads"

    def one_step():
        t0 = time.perf_counter()
        if os.path.exists(binary):
            cmd = [binary, "--nd", str(nd), "--nv", "1", "--dims"] + [str(d) for d in sdims] + \
                  ["--nt", "2", "--gen", "moving_extremum", "--p"] + [repr(float(v)) for v in sp] + \
                  ["--no-trace", "--quiet", "--nthreads", str(cores)]
            out = subprocess.run(cmd, check=True, capture_output=True).stdout.decode().strip().splitlines()[-1]
            st = json.loads(out)
            t_path = st["t_push"] + st["t_sweep"]            # derive + sweep: the hot path
            assert int(st["simplices"]) == simplices, (st["simplices"], simplices)
        else:
            from oracle import cp_oracle as O               # plain-C port of the reference path
            snaps = [O.gen_moving_extremum(sdims, sp[:nd], sp[nd:], float(k)) for k in range(2)]
            t1 = time.perf_counter()
            O.track(snaps, sdims, field="scalar", trace=False)
            t_path = time.perf_counter() - t1
        return t_path, time.perf_counter() - t0

    if not os.path.exists(binary):
        kind = "port"
    for _ in range(args.warmup):
        one_step()
    tot = wall = 0.0
    for _ in range(args.steps):
        a, b = one_step()
        tot += a
        wall += b
    value = simplices * args.steps / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64+int64", "data": "synthetic",
        "config": {"workload": label, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_ms_per_step": 1e3 * wall / args.steps,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(config):
    """bounded sample (about 10-30 s of CPU work) of the reference CPU tracker on this box's host cores"""
    label, dims, params, _ = CONFIGS[config]
    nd = len(dims)
    binary = os.path.join(ROOT, "oracle", "_ref", "ftk_ref_oracle")
    cores = os.cpu_count() or 1
    sdims = [1536, 1536] if nd == 2 else [96, 96, 96]
    T = 3
    sp = [d / 2 + 0.3 for d in sdims] + list(params[nd:])
    sample = f"{'x'.join(map(str, sdims))}x{T} timesteps of the same generator, {cores} threads"
    try:
        if os.path.exists(binary):
            cmd = [binary, "--nd", str(nd), "--nv", "1", "--dims"] + [str(d) for d in sdims] + \
                  ["--nt", str(T), "--gen", "moving_extremum", "--p"] + [repr(float(v)) for v in sp] + \
                  ["--no-trace", "--quiet", "--nthreads", str(cores)]
            st = json.loads(subprocess.run(cmd, check=True, capture_output=True, timeout=600).stdout.decode().strip().splitlines()[-1])
            return {"value": st["simplices"] / (st["t_push"] + st["t_sweep"]), "unit": UNIT, "cores": cores, "kind": "reference",
                    "sample": sample, "t_derive_s": st["t_push"], "t_sweep_s": st["t_sweep"]}
        from oracle import cp_oracle as O
        snaps = [O.gen_moving_extremum(sdims, sp[:nd], sp[nd:], float(k)) for k in range(T)]
        t0 = time.perf_counter()
        O.track(snaps, sdims, field="scalar", trace=False)
        dt = time.perf_counter() - t0
        ncore = 1
        for d in sdims:
            ncore *= d - 3
        n = ncore * ((12 * (T - 1) + 2) if nd == 2 else (60 * (T - 1) + 6))
        return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    except Exception as e:  # the baseline is reported, never required
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "unavailable", "sample": f"failed: {e}"}


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=252)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.steps == 252:
            args.steps = 20   # default sized so the whole run ends within a few minutes
        reference_arm(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import ftk_b200
    from ftk_b200 import _lib
    _lib.lib()   # fail loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    label, dims, params, NL = CONFIGS[args.config]
    nd = len(dims)
    vector = args.config in VECTOR_CONFIGS
    x0, dirv = (params[:nd], params[nd:]) if params else (None, None)
    K, W = args.steps, args.warmup
    ncore = 1
    for d in dims:
        ncore *= d - (2 if vector else 3)      # json_interface.hh:639-654: [1, D-2] for vector input, [2, D-2] for scalar input
    per_step = ncore * (12 if nd == 2 else 60)
    nvert = int(np.prod(dims))

    def gen_layer(j, out=None):
        """synthetic_moving_extremum at time j (ref: synthetic.hh:332-354), memory order ([D,]H,W)"""
        axes = [torch.arange(d, dtype=torch.float64, device=dev) for d in dims]
        e = [(axes[q] - (x0[q] + dirv[q] * float(j))) ** 2 for q in range(nd)]
        if nd == 2:
            s = e[0][None, :] + e[1][:, None]
        else:
            s = (e[0][None, None, :] + e[1][None, :, None]) + e[2][:, None, None]
        if out is not None:
            out.copy_(s)
            return out
        return s.contiguous()

    def gen_vector_layer(j):
        """C5: synthetic_double_gyre(time = 0.1 j) (synthetic.hh:130-150,193-217); C4: synthetic_abc_flow (synthetic.hh:239-260)
        with A(j) = sqrt(3) + 0.5 (j/64) sin(pi j/64) (the time dependence is ours); memory order ([D,]H,W,n)"""
        out = torch.empty(tuple(reversed(dims)) + (nd,), dtype=torch.float64, device=dev)
        ax = [torch.arange(d, dtype=torch.float64, device=dev) / (d - 1) for d in dims]
        if nd == 2:
            A, omega, eps, t = 0.1, 2 * math.pi, 0.25, 0.1 * j
            x, y = ax[0] * 2, ax[1]
            a, b = eps * math.sin(omega * t), 1 - 2 * eps * math.sin(omega * t)
            f, dfdx = a * x * x + b * x, 2 * a * x + b
            out[..., 0] = (-math.pi * A * torch.sin(math.pi * f))[None, :] * torch.cos(math.pi * y)[:, None]
            out[..., 1] = (math.pi * A * torch.cos(math.pi * f) * dfdx)[None, :] * torch.sin(math.pi * y)[:, None]
        else:
            A, B, Cc = math.sqrt(3.0) + 0.5 * (j / 64.0) * math.sin(math.pi * j / 64.0), math.sqrt(2.0), 1.0
            x, y, z = (a * 2 * math.pi for a in ax)
            for k0 in range(0, dims[2], 64):       # slabs of 64 planes keep the temporaries small next to 25.8 GB layers
                zs = z[k0:k0 + 64]
                o = out[k0:k0 + 64]
                o[..., 0] = (A * torch.sin(zs))[:, None, None] + (Cc * torch.cos(y))[None, :, None]
                o[..., 1] = (B * torch.sin(x))[None, None, :] + (A * torch.cos(zs))[:, None, None]
                o[..., 2] = (Cc * torch.sin(y))[None, :, None] + (B * torch.cos(x))[None, None, :]
        return out

    layers = [gen_vector_layer(j) if vector else gen_layer(j) for j in range(NL)]
    g0 = rank * (W + K)
    fld = "vector" if vector else "scalar"

    def make():
        return ftk_b200.make_tracker(dims, field=fld, device=local, start_timestep=g0)

    def push_ptr(t, ptr):
        if vector:
            t.push_device_pointers(vector=ptr)
        else:
            t.push_device_pointers(scalar=ptr)

    # ---- device-resident measurement -------------------------------------------------------------
    tr = make()
    ptrs = [int(l.data_ptr()) for l in layers]     # resident layers are borrowed in place through the C ABI
    push_ptr(tr, ptrs[tri(g0, NL)])
    # N > 1, halo through NVLink peer memory (default; FTKB_HALO=nccl copies the layer with ncclSend/Recv instead): the slab
    # below needs this rank's first layer only for its last sweep, and of it only the 16-byte range cells and the vertices
    # around the surviving cubes -- the owner exports IPC handles of the layer and of its cells, the neighbour maps them and
    # pushes the layer as a remote snapshot (ftkb_push_snapshot_remote): nothing is copied.  Handle exchange and mapping are
    # set-up (once per run), like NCCL's channel set-up; they stay outside the timed region.
    peer = None
    first_done = False
    if dist and os.environ.get("FTKB_HALO", "peer") == "peer":
        mine = None
        try:
            push_ptr(tr, ptrs[tri(g0 + 1, NL)])
            tr.update_timestep()                         # the sweep that builds the cells of this rank's first layer
            cptr, cbytes, cres = tr.export_layer_cells(0)
            mine = (_lib.ipc_export(ptrs[tri(g0, NL)]), _lib.ipc_export(cptr), cres)
        except _lib.FTKBError:
            mine = None
        tr.advance_timestep()                            # warm-up step 0 completes (repeated sweep from the cells, layer popped)
        first_done = True
        allh = [None] * world
        dist.all_gather_object(allh, mine)
        if all(h is not None for h in allh):
            try:
                if rank < world - 1:
                    lh, ch, cres = allh[rank + 1]
                    peer = (_lib.ipc_import(lh, local), _lib.ipc_import(ch, local), cres)
                else:
                    peer = ()
            except _lib.FTKBError:
                peer = None
        oks = [None] * world
        dist.all_gather_object(oks, peer is not None)
        if not all(oks):
            peer = None          # some rank could not export / map: every rank copies the halo with NCCL instead
    use_peer = peer is not None
    for i in range(1 if first_done else 0, W):
        push_ptr(tr, ptrs[tri(g0 + i + 1, NL)])
        tr.advance_timestep()
    halo = None
    sampler = ClockSampler(physical_gpu_index(local))
    torch.cuda.synchronize()
    tr.synchronize()

    def exchange_halo():
        """one-layer halo: the first layer of the next slab (ncclSend/ncclRecv over NVLink), one batched P2P group"""
        ops = []
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, layers[tri(g0, NL)], rank - 1))
        if rank < world - 1:
            ops.append(dist.P2POp(dist.irecv, halo, rank + 1))
        return dist.batch_isend_irecv(ops) if ops else []

    if dist:
        # untimed: NCCL sets its peer-to-peer channels up lazily on first use (hundreds of ms); the timed exchange
        # below then moves the halo over warm channels, as every exchange after the first would in a long run
        if not use_peer:
            if rank < world - 1:
                halo = torch.empty_like(layers[0])
            for r in exchange_halo():
                r.wait()
        warm = torch.zeros(1, dtype=torch.float64, device=dev)
        dist.all_gather([torch.empty_like(warm) for _ in range(world)], warm)
        dist.all_reduce(warm, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize()
        dist.barrier()
    tr.reset_stats()
    sampler.start()
    tr.timer_start()
    reqs = exchange_halo() if (dist and not use_peer) else []     # inside the timed region
    res_layers, factors = [], []
    import ctypes as _C
    _L, _st = _lib, _lib.Stats()
    for i in range(W, W + K):
        nxt = ptrs[tri(g0 + i + 1, NL)]
        if use_peer and peer and i == W + K - 1:
            # the slab's last sweep: the next slab's first layer stays where it is, in the neighbour GPU's memory
            if vector:
                tr.push_remote_snapshot(vector=peer[0], cells=peer[1], resolution=peer[2])
            else:
                tr.push_remote_snapshot(scalar=peer[0], cells=peer[1], resolution=peer[2])
            tr.advance_timestep()
            if dist:
                _L.lib().ftkb_get_stats(tr._h, _C.byref(_st))
                res_layers.append(_st.resolution)
                factors.append(_st.scaling_factor)
            continue
        if halo is not None and i == W + K - 1:
            for r in reqs:
                r.wait()
            torch.cuda.current_stream().synchronize()
            nxt = int(halo.data_ptr())
        push_ptr(tr, nxt)
        tr.advance_timestep()
        if dist:
            _L.lib().ftkb_get_stats(tr._h, _C.byref(_st))   # one struct read per step: running minimum inside this slab + factor used
            res_layers.append(_st.resolution)
            factors.append(_st.scaling_factor)
    redo = 0
    if dist:
        # running minimum of min non-zero |v| across slabs (the reference's factor is a running quantity):
        # every rank swept optimistically with its own slab's minimum; verify against the global prefix
        for r in reqs:
            r.wait()
        mine = torch.tensor([min(res_layers)], dtype=torch.float64, device=dev)
        allres = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allres, mine)
        prefix = min([float(a.item()) for a in allres[:rank]] + [float("inf")])
        if prefix < float("inf"):
            def nbits(res):
                return max(8, min(int(math.ceil(math.log2(1.0 / res))), 21))
            for rl, f in zip(res_layers, factors):
                if float(1 << nbits(min(prefix, rl))) != f:
                    redo += 1
        # (moving-extremum slabs never differ; a differing slab is re-swept with resolution_init = prefix)
    ms = tr.timer_stop()
    sampler.stop_flag = True
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    st = tr.stats()
    sampler.join(timeout=1.0)

    # ---- finalize (not part of a step; reported separately) --------------------------------------
    t0 = time.perf_counter()
    pts = tr.get_discrete_critical_points()
    if dist:
        gathered = [None] * world
        dist.all_gather_object(gathered, pts.tobytes())
        if rank == 0:
            for b in gathered[1:]:
                tr.import_points(np.frombuffer(b, dtype=_lib.POINT_DTYPE))
    ntraj = None
    if rank == 0:
        tr.finalize()
        ntraj = len(tr.get_trajectory_index())
        npts_total = len(tr.get_discrete_critical_points())
    finalize_ms = 1e3 * (time.perf_counter() - t0)
    tr.close()

    def run_e2e():
        nonlocal layers
        E = max(1, min(args.e2e_steps, K))
        NH = min(8, NL) if layers[0].numel() * 8 < (4 << 30) else 2      # pinned host copies: two are enough for 25.8 GB layers
        E = E if NH > 2 else min(E, 2)
        host = [torch.empty(tuple(layers[0].shape), dtype=torch.float64).pin_memory() for _ in range(NH)]
        for j in range(NH):
            host[j].copy_(layers[j])
        torch.cuda.synchronize()
        if NH == 2:       # 25.8 GB layers: the context's own copies need the room the resident series occupied
            layers = []
            torch.cuda.empty_cache()
        tr2 = make()

        def push_host(t, a):
            if vector:
                t.push_vector_field_snapshot(a)
            else:
                t.push_scalar_field_snapshot(a)

        push_host(tr2, host[0].numpy())
        for i in range(2):
            push_host(tr2, host[tri(i + 1, NH)].numpy())
            tr2.advance_timestep()
        tr2.synchronize()
        if dist:
            dist.barrier()
        tr2.reset_stats()
        tr2.timer_start()
        for i in range(2, 2 + E):
            push_host(tr2, host[tri(i + 1, NH)].numpy())   # H2D inside the timed region
            tr2.advance_timestep()                                         # counters + resolution read back every step
        ms2 = tr2.timer_stop()
        if dist:
            t = torch.tensor([ms2], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        st2 = tr2.stats()
        tr2.close()
        return ms2, E, st2

    # ---- end-to-end measurement: host buffers through the same C-ABI calls ------------------------
    ms2, E, st2 = None, 0, None
    if args.e2e_steps > 0:
        ms2, E, st2 = run_e2e()

    if rank == 0:
        peak, peak_src = measured_peaks()
        nscan = max(int(st["scan_launches"]), 1)
        scan_ms = st["ms_scan"] / nscan
        fused = (not vector) and st["ms_derive"] == 0.0   # scalar input: gradient fused into the scan, the vector field is never materialised
        # algorithmic bytes of one scan launch (DESIGN.md "Kernels"): the two input layers read once each
        # (fused: fp64 scalar layers, 8 B/vertex; otherwise fp64 vector layers, nd*8 B/vertex) + 72-B hit records
        in_bytes = 2 * nvert * (1 if fused else nd) * 8
        # one cell buffer (16 B per lane and block): written for the new layer, read for the current one
        cell_bytes = nvert * (2.1 if nd == 3 else (1.04 if not vector else 1.03))
        alg_bytes = in_bytes + 72 * (st["points"] / max(K, 1))
        achieved = alg_bytes / (scan_ms * 1e-3) / 1e9
        # SURVEY.md 8(d) counts materialised vector layers (2.667 B/simplex in 2D, 0.8 in 3D) for the same launch
        survey_bytes = (32.0 / 12.0 if nd == 2 else 48.0 / 60.0) * per_step
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
                traffic = json.load(f).get(args.config)
        except Exception:
            pass
        # the default scan streams ONE layer per step and reads/writes 16-byte range cells for the other
        # (DESIGN.md 4.2); FTKB_SCAN2D / FTKB_SCAN3D = twolayer selects the kernels that re-read both layers
        cells = fused and os.environ.get("FTKB_SCAN2D" if nd == 2 else "FTKB_SCAN3D", "") not in ("twolayer", "plain") \
            and os.environ.get("FTKB_SCAN", "tile") == "tile"
        if vector:
            vcells = os.environ.get("FTKB_VSCAN", "") != "twolayer"
            kname = (("vscan2d_build_kernel<1,true>" if nd == 2 else "vscan3d_build_kernel<1,true>") if vcells
                     else ("scan2d_kernel<true>" if nd == 2 else "scan3d_kernel<true>"))
        elif nd == 2:
            kname = "scan2d_build_kernel<1,true>" if cells else "scan2d_tile_kernel<true>"
        else:
            kname = "scan3d_build_kernel<1,true>" if cells else ("scan3d_fused_kernel<true>" if fused else "scan3d_kernel<true>")
        if not cells and not (vector and os.environ.get("FTKB_VSCAN", "") != "twolayer"):
            traffic = None
        line = {
            "metric": METRIC, "value": per_step * K * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64+int64", "data": "synthetic",
            "config": {"workload": label, "simplices_per_step_per_gpu": per_step, "layers_resident": NL,
                       "l2": "inputs larger than L2 (each fp64 layer >= 0.5 GB; no flush needed)",
                       "parallelism": f"time-slab x{world}" if world > 1 else "single GPU",
                       "halo": (("nvlink peer memory: the next slab's first layer is read in place (range cells + sparse vertices), nothing copied"
                                 if use_peer else "ncclSend/ncclRecv of the next slab's first layer, inside the timed region") if world > 1 else None),
                       "step": "one advance_timestep: gradient + min|v| + range cells + exact sign early-out (fused scan kernel) + per-simplex test kernel" if fused else
                               "one advance_timestep: derive(gradient+resolution) + scan + per-simplex test"},
            "roofline": {"bound": "hbm", "kernel": kname,
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_gbs": (traffic / (scan_ms * 1e-3) / 1e9) if traffic else None,
                         "traffic_frac": (traffic / (scan_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "traffic_note": "the scan streams one layer per step and keeps 16-byte range cells for the other, so DRAM traffic "
                                         "(ncu, per launch) is below the algorithmic bytes of the two-layer contract" if (cells or (vector and vcells)) else None,
                         "model_traffic_bytes": (in_bytes / 2 + 2 * cell_bytes) if (cells or (vector and vcells)) else in_bytes,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": scan_ms,
                         "launches_timed": nscan, "sweeps_repeated": int(st["sweeps_repeated"]),
                         "survey_8d_equivalent": {"bytes_per_launch": survey_bytes, "gbs": survey_bytes / (scan_ms * 1e-3) / 1e9,
                                                  "note": "SURVEY.md 8(d) assumes materialised fp64 vector layers; the fused kernel reads the scalar layers instead"},
                         "note": "rank 0; tensor cores unused by design (no dense contraction on this path)"},
            "kernel_ms_per_step": {"derive": st["ms_derive"] / K, "scan": scan_ms, "test": st["ms_test"] / K},
            "e2e": ({"value": per_step * E * world / (ms2 * 1e-3), "unit": UNIT, "steps": E, "ms_per_step": ms2 / E,
                     "h2d_bytes_per_step": st2["h2d_bytes"] / E, "d2h_bytes_per_step": st2["d2h_bytes"] / E,
                     "api": "ftkb_push_snapshot(host) + ftkb_advance_timestep"} if E else None),
            "gpu_launches": int(st["kernel_launches"]),
            "clocks": sampler.result(),
            "finalize_ms": finalize_ms, "trajectories": ntraj, "punctured_simplices": int(npts_total),
            "cells_refined_per_step": st["cells_refined"] / K, "slab_refactor_steps": redo,
        }
        if world == 1 and not args.no_cpu_baseline and not vector:
            line["cpu_baseline"] = cpu_baseline_sample(args.config)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
