#!/usr/bin/env python
"""bench.py -- QGMAP pixel-iterations/s on B200 (contract: the task statement; design notes in DESIGN.md section 5).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA kernels through the C ABI of libqgmap.so)
  python bench.py --impl reference --gpus N --steps K ...   reference arm: the CPU fp64 restatement on the host cores

BASELINE.json quotes its metric on "640x480 & 4K synthetic at 1/2/4/8 B200"; both are measured, every run:
  value  : ONE synthetic 3840x2160 frame pair, L=3 mixture components, K=5 (5x5 Gauss-Hermite) -- BASELINE configs[3].  At N GPUs
           the frame is split into N row bands (one rank each); per iteration the boundary rows travel between neighbours and 4L
           doubles are summed over all ranks, by libqgmap's own kernels over NVLink peer memory (strong scaling).
  extra  : a batch of synthetic 640x480 pairs, L=3, K=5, sharded by pair -- BASELINE configs[4] (256 pairs over 8 GPUs = 32 per
           GPU; every rank runs 32, no communication: weak scaling).  Reported under "batch640" in the same JSON line.
  e2e    : the 4K pair through the reference-facing call with HOST buffers: gqmap_gpu_mixture(options,I1,I2) at N=1, its
           multi-process twin dist.gqmap_gpu_mixture_bands at N>1 (frames and initial state from pinned host memory, all iterations
           with the reference's monitoring cadence, beliefs and histories back to the host).
One STEP = `iters` ascent iterations of the frame pair (chosen so that the K timed steps last >= ~3.5 s; printed in config).
pixel-iterations/s = image pixels x iterations / seconds.
"""
import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "gqmap-opticalflow_b200"

L_MIX, K_GH = 3, 5
M_4K, N_4K = 2160, 3840
WORKLOAD = ("one synthetic 3840x2160 frame pair (integer grey levels 0..255 like the reference's rgb2gray frames), L=3 mixture components, "
            "K=5 (5x5 Gauss-Hermite), gqmap_gpu_mixture path (BASELINE configs[3])")


def flops_per_px_it(L, K, super_=False):
    """algorithmic FP32 work per pixel-iteration (the contract for roofline.achieved)."""
    return (102.25 * L * K * K + 15 * L) if super_ else (270.0 * L * K * K + 240.0 * L)


def bytes_per_px_it(L, super_=False):
    return (72.0 * L / 16 + 8) if super_ else (72.0 * L + 8)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def options_for(extrema, **extra):
    """optical_flow.m:16-23 with L=3, K=5 and the clamp range of the synthetic ground truth."""
    minu, maxu, minv, maxv = extrema
    o = dict(K=K_GH, L=L_MIX, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv,
             maxv=maxv)
    o.update(extra)
    return o


def init_arrays(opts, M, N, L, seed):
    """gqmap_gpu_mixture.m:18-24 with NumPy's generator."""
    rng = np.random.default_rng(seed)
    w = rng.random(L)
    f = lambda a: np.asfortranarray(a)
    return dict(w=w, muu=f(opts["minu"] + rng.random((M, N, L)) * (opts["maxu"] - opts["minu"])),
                muv=f(opts["minv"] + rng.random((M, N, L)) * (opts["maxv"] - opts["minv"])),
                sigmau=f(rng.random((M, N, L)) + (opts["maxu"] - opts["minu"])),
                sigmav=f(rng.random((M, N, L)) + (opts["maxv"] - opts["minv"])),
                pn=f(np.zeros((M, N, L))), rou=f(np.zeros((M, N, L, 2, 2))))


def pinned_like(torch, a):
    """Copy a NumPy array into page-locked host memory, keeping MATLAB (column-major) layout."""
    t = torch.empty(a.size, dtype=torch.float64).pin_memory()
    v = t.numpy().reshape(a.shape, order="F")
    v[...] = a
    return v, t


def shared_4k_pair(pkg, torch, dist, rank, world, M, N):
    """The 4K pair is generated once (rank 0, ~20 s of host time) and broadcast; every rank needs the whole frames."""
    if world == 1:
        I1, I2, flow, ext = pkg.synthetic_pair(M, N, seed=1234, grey_levels=True)
        return I1, I2, ext
    buf = torch.empty(2 * M * N + 4, dtype=torch.float64, device="cuda")
    if rank == 0:
        I1, I2, flow, ext = pkg.synthetic_pair(M, N, seed=1234, grey_levels=True)
        buf.copy_(torch.from_numpy(np.concatenate([I1.ravel(order="F"), I2.ravel(order="F"), np.asarray(ext)])))
    dist.broadcast(buf, src=0)
    h = buf.cpu().numpy()
    I1 = np.asfortranarray(h[:M * N].reshape((M, N), order="F"))
    I2 = np.asfortranarray(h[M * N:2 * M * N].reshape((M, N), order="F"))
    return I1, I2, tuple(float(v) for v in h[2 * M * N:])


def _gen_pair(seed):
    sys.path.insert(0, os.path.join(ROOT, PKG))
    import frames
    return frames.synthetic_pair(480, 640, seed=seed, grey_levels=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the QGMAP path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG)
    M, N, L, K = args.band_rows, args.band_cols, L_MIX, K_GH

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxf(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    I1, I2, ext = shared_4k_pair(pkg, torch, dist, rank, world, M, N)
    opts = options_for(ext, device=local)

    # ---------------- device-resident leg: `value` (one 4K pair in `world` row bands) ----------------
    bopts = dict(opts)
    if world > 1:
        rb, re = pkg.dist.band_rows(M, rank, world)
        bopts.update(row_begin=rb, row_end=re)
    s = pkg.Solver(bopts, I1, I2)
    transport = "none"
    if world > 1:
        transport = pkg.dist.connect_band(s, dist, device=dev, transport=args.band_transport)
    s.init_state(seed=4321)
    barrier()
    r = s.step(args.init_iters)                                       # the random-init phase, timed on its own (roofline.init_phase)
    init_ms = maxf(r["ms"]) / max(r["n_done"], 1)
    if args.burnin > args.init_iters:
        s.step(args.burnin - args.init_iters)
    probe = s.step(50)                                                # sizes a step: K timed steps should last >= --min-seconds
    ms_it = maxf(probe["ms"]) / 50
    iters = args.iters if args.iters > 0 else max(50, int(math.ceil(args.min_seconds * 1e3 / (ms_it * max(args.steps, 1)) / 50.0)) * 50)
    for _ in range(args.warmup):
        s.step(iters)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ms, launches = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = s.step(iters)
        ms += r["ms"]
        launches += r["launches"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    s.close()
    ms_max = maxf(ms)
    value = M * N * iters * args.steps / (ms_max * 1e-3)
    kern_ms = ms_max / (iters * args.steps)                           # one launch == one iteration of this rank's band

    # ---------------- extra: batch of 640x480 pairs, sharded by pair (BASELINE configs[4]: 32 per GPU) ----------------
    batch = None
    if args.batch_pairs > 0:
        from multiprocessing import get_context
        seeds = [1234 + 2 * (rank * args.batch_pairs + p_) for p_ in range(args.batch_pairs)]
        with get_context("spawn").Pool(max(1, min(args.batch_pairs, (os.cpu_count() or 8) // world))) as pool:
            frames_ = pool.map(_gen_pair, seeds)
        solvers = []
        for p_, (J1, J2, flow, e_) in enumerate(frames_):
            sv = pkg.Solver(options_for(e_, device=local), J1, J2)
            sv.init_state(seed=4321 + p_)
            solvers.append(sv)
        del frames_
        b_init, _ = pkg.batch_step(solvers, 100)
        b_init = maxf(b_init) / 100
        pkg.batch_step(solvers, max(args.batch_burnin - 100, 0))
        b_probe, _ = pkg.batch_step(solvers, 20)
        b_ms_it = maxf(b_probe) / 20
        b_iters = max(20, int(math.ceil(args.min_seconds * 1e3 / (b_ms_it * max(args.steps, 1)) / 10.0)) * 10)
        for _ in range(min(args.warmup, 2)):
            pkg.batch_step(solvers, b_iters)
        bs = ClockSampler(local)
        bs.start()
        barrier()
        b_ms, b_launch = 0.0, 0
        for _ in range(args.steps):
            m_, nl = pkg.batch_step(solvers, b_iters)
            b_ms += m_
            b_launch += nl
        barrier()
        b_clocks = bs.stop()
        for sv in solvers:
            sv.close()
        b_ms_max = maxf(b_ms)
        b_value = world * args.batch_pairs * 480 * 640 * b_iters * args.steps / (b_ms_max * 1e-3)
        batch = dict(ms=b_ms_max, iters=b_iters, launches=b_launch, value=b_value, clocks=b_clocks, init_ms_it=b_init,
                     burnin=args.batch_burnin)

    # ---------------- end-to-end leg: the public call with HOST buffers ----------------
    # options.its: the reference drivers' 30000 unless that would exceed --e2e-budget seconds at the measured iteration time
    its_e2e = max(300, min(args.e2e_its, int(args.e2e_budget * 1e3 / ms_it / 300) * 300))
    rows_ = np.arange(M, dtype=np.float64).reshape(M, 1)
    cols_ = np.arange(N, dtype=np.float64).reshape(1, N)
    tflow = np.asfortranarray(np.stack([3.0 * np.sin(2 * np.pi * rows_ / M) + 1.0 + 0.0 * cols_,
                                        2.0 * np.cos(2 * np.pi * cols_ / N) + 0.0 * rows_], axis=2))   # frames.synthetic_pair's ground truth
    unk = np.zeros((M, N), dtype=np.uint8, order="F")
    init = init_arrays(opts, M, N, L, 999)
    pin, keep, h2d = {}, [], 0
    for k_, a in list(init.items()) + [("I1", I1), ("I2", I2), ("tflow", tflow)]:
        pin[k_], t_ = pinned_like(torch, a)
        keep.append(t_)
    del init
    e2e_opts = dict(opts, its=its_e2e, init={k_: pin[k_] for k_ in ("muu", "muv", "sigmau", "sigmav", "pn", "rou", "w")},
                    trueFlow=pin["tflow"], unknownIdx=unk)
    if world > 1:
        rb, re = pkg.dist.band_rows(M, rank, world)
        rows_mine = (min(re + 1, M) - max(rb - 1, 0))
    else:
        rows_mine = M
    h2d = 2 * M * N * 8 + 9 * L * rows_mine * N * 8 + 2 * M * N * 8 + M * N   # frames, this rank's rows of the 9L state planes, trueFlow, mask
    d2h = (4 * L * (rows_mine if world == 1 else (re - rb)) * N + L + its_e2e) * 8      # mu, sigma (own rows), alpha, Energy
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        if world > 1:
            out = pkg.dist.gqmap_gpu_mixture_bands(e2e_opts, pin["I1"], pin["I2"], dist, device=dev, transport=args.band_transport)
        else:
            out = pkg.gqmap_gpu_mixture(e2e_opts, pin["I1"], pin["I2"])
    barrier()
    e2e_s = maxf(time.perf_counter() - t0)
    its_done = int(np.count_nonzero(out[4])) or its_e2e               # Energy(it) is 0 past an early stop (:16)
    e2e_value = M * N * its_done * args.e2e_steps / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fp32_meas = pkg.fp32_peak(local)
    fp32_nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    F, B = flops_per_px_it(L, K), bytes_per_px_it(L)
    px_launch = M * N / world                                          # pixels one launch (one band) processes
    ach_tf = px_launch * F / (kern_ms * 1e-3) / 1e12
    ach_gbs = px_launch * B / (kern_ms * 1e-3) / 1e9
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    if world > 1 and prof:                                             # a band: the one-of-eight capture at N=8, else the whole-frame capture scaled
        band = prof.get("band_of_8", {})
        if world == 8 and band:
            prof = dict(prof, **band)
        else:
            prof = dict(prof, dram_bytes_per_launch=prof.get("dram_bytes_per_launch", 0) / world,
                        source=prof.get("source", "") + " / %d (whole-frame capture scaled to one band)" % world)
    out_line = {
        "metric": "QGMAP pixel-iterations/s", "value": value, "unit": "pixel-iter/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pixels": M * N, "L": L, "K": K, "iters_per_step": iters,
                   "parallelism": ("row bands x%d, 1-row halo exchange + 4L-double global sum per iteration, transport %s (p2p = stores and "
                                   "flags over NVLink peer memory from inside the iteration kernel; nccl = send/recv + all-reduce)" % (world, transport))
                                  if world > 1 else "1 GPU, undivided frame",
                   "trajectory_window": "timed iterations %d..%d of the ascent from the reference's random init (gqmap_gpu_mixture.m:18-24); "
                                        "roofline.init_phase is the first %d iterations" % (
                                            args.burnin + 50 + args.warmup * iters + 1, args.burnin + 50 + (args.warmup + args.steps) * iters,
                                            args.init_iters),
                   "l2": "no explicit flush: state 2 x %.0f MB + gather layout %.0f MB per GPU exceed the 126 MB L2 and are streamed every "
                         "iteration" % (M * N * 9 * L * 4 / 1e6 / world, (M + 2) * N * 32 / 1e6 / world)},
        "clocks": clocks, "gpu_launches": int(launches), "wall_s_timed": wall,
        "e2e": {"value": e2e_value, "unit": "pixel-iter/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "its_per_call": its_e2e, "its_done": its_done, "seconds_per_call": e2e_s / args.e2e_steps, "steps": args.e2e_steps,
                "api": ("gqmap_gpu_mixture(options,I1,I2) -> qgmap_solve (C ABI)" if world == 1 else
                        "dist.gqmap_gpu_mixture_bands(options,I1,I2,...) -> qgmap_create/set_state/step/monitor_partial/get_state (C ABI), one rank per band")
                       + ": pinned host frames and initial state in, host arrays out, monitoring (MAP, AEPE, logP) at it=1 and every 300 as "
                         "the reference (gqmap_gpu_mixture.m:52-68); options.its = %d (the reference drivers use 30000, optical_flow.m:17)" % its_e2e},
        "roofline": {"bound": "fp32", "achieved": ach_tf, "peak": fp32_meas, "unit": "TFLOP/s", "frac": ach_tf / fp32_meas,
                     "peak_source": "FFMA micro-benchmark measured live on this GPU (qgmap_fp32_peak); MEASURED_PEAKS.json holds no FP32 figure; "
                                    "nominal 148x128x2x1.965GHz = %.1f" % fp32_nominal,
                     "frac_of_nominal": ach_tf / fp32_nominal,
                     "kernel": "qgmap_iter_kernel<5,false,false> on this rank's band (%d px), %.4f ms per launch (CUDA events on its stream, "
                               "exchange included at N>1)" % (int(px_launch), kern_ms),
                     "algorithmic": "%.0f flop and %.0f bytes per pixel-iteration (SURVEY 8d); the kernel executes fewer instructions than "
                                    "this count (moments, affine edges), so frac is not pipe utilisation: see fma_pipe_active" % (F, B),
                     "init_phase": {"ms_per_launch": init_ms, "frac": px_launch * F / (init_ms * 1e-3) / 1e12 / fp32_meas,
                                    "iterations": "1..%d" % args.init_iters},
                     "fma_pipe_active": prof.get("fma_pipe_active_pct"), "issue_active": prof.get("issue_active_pct"),
                     "l1_data_stage_active": prof.get("l1_data_stage_active_pct"),
                     "traffic": prof.get("dram_bytes_per_launch"), "traffic_source": prof.get("source"),
                     "hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}},
    }
    if batch:
        bv = batch["value"]
        b_kern_tf = bv * F / 1e12 / world
        out_line["batch640"] = {
            "workload": "batch of synthetic 640x480 frame pairs (seeds 1234+2p), L=3, K=5, gqmap_gpu_mixture path, %d pairs per GPU "
                        "(BASELINE configs[4]: 256 pairs over 8 GPUs), sharded by pair, no communication" % args.batch_pairs,
            "value": bv, "unit": "pixel-iter/s", "scaling": "weak", "n_gpus": world, "pairs_per_gpu": args.batch_pairs,
            "iters_per_step": batch["iters"], "ms_per_step": batch["ms"] / args.steps, "gpu_launches": int(batch["launches"]),
            "clocks": batch["clocks"],
            "trajectory_window": "after %d iterations per pair from the random init" % batch["burnin"],
            "roofline": {"bound": "fp32", "achieved": b_kern_tf, "peak": fp32_meas, "unit": "TFLOP/s", "frac": b_kern_tf / fp32_meas,
                         "note": "per GPU, whole batch (pairs advance concurrently on their own streams)",
                         "init_phase_frac": args.batch_pairs * 480 * 640 * F / (batch["init_ms_it"] * 1e-3) / 1e12 / fp32_meas}}
    if not args.no_cpu and world == 1:                 # reported at N=1 only (rank 0's host cores)
        out_line["cpu_baseline"] = cpu_baseline(I1, I2, ext, budget_s=args.cpu_budget, nthreads=os.cpu_count() or 1)
    print(json.dumps(out_line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(I1, I2, ext, budget_s=20.0, nthreads=0):
    """The oracle (kind 'port': fp64 C restatement of gqmap_gpu_mixture.m, OpenMP) timed on the host cores on a bounded
    sample of the same workload: whole iterations of the 4K pair, as many as fit the budget."""
    from oracle import oracle as O
    M, N = I1.shape
    minu, maxu, minv, maxv = ext
    cfg = O.make_config(M, N, L_MIX, K_GH, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, nthreads=nthreads)
    st = O.init_state(cfg, 4321)
    VV = O.get_vv(I2)
    t0 = time.perf_counter()
    O.run(cfg, I1, VV, st, 1, 10 ** 6, 1)
    t1 = time.perf_counter() - t0
    n = max(1, min(20, int(budget_s / max(t1, 1e-3)) - 1))
    t0 = time.perf_counter()
    O.run(cfg, I1, VV, st, 2, 10 ** 6, n)
    dt = time.perf_counter() - t0
    cores = os.cpu_count() if nthreads == 0 else nthreads
    return {"value": M * N * n / dt, "unit": "pixel-iter/s", "cores": cores, "kind": "port",
            "sample": "%d whole iteration(s) of the %dx%d pair, L=%d K=%d, fp64 C restatement of gqmap_gpu_mixture.m with OpenMP "
                      "(the MATLAB reference cannot run here)" % (n, N, M, L_MIX, K_GH)}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path = the fp64 oracle port, all host threads, a bounded
    sample of the same workload per step (--ref-iters whole iterations of the 4K pair)."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, PKG))
    import frames                                       # workload generation only (no libqgmap, no GPU)
    from oracle import oracle as O
    M, N = args.band_rows, args.band_cols
    I1, I2, flow, (minu, maxu, minv, maxv) = frames.synthetic_pair(M, N, seed=1234, grey_levels=True)
    cores = os.cpu_count() or 1
    cfg = O.make_config(M, N, L_MIX, K_GH, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, nthreads=cores)   # torchrun exports OMP_NUM_THREADS=1
    VV = O.get_vv(I2)
    st = O.init_state(cfg, 4321)
    it, iters = 1, args.ref_iters

    def step():
        nonlocal it
        O.run(cfg, I1, VV, st, it, 10 ** 6, iters)
        it += iters
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = M * N * iters * args.steps / dt
    sample = ("%d whole iteration(s) of the 4K pair per step (%d px), fp64 C restatement of gqmap_gpu_mixture.m (oracle port; MATLAB/Octave "
              "absent), OpenMP on all %d host threads" % (iters, M * N, cores))
    out = {"impl": "reference", "metric": "QGMAP pixel-iterations/s", "value": v, "unit": "pixel-iter/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "pixels": M * N, "L": L_MIX, "K": K_GH, "iters_per_step": iters,
                      "parallelism": "host cores (OpenMP), no GPU"},
           "cpu_baseline": {"value": v, "unit": "pixel-iter/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "pixel-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=0, help="ascent iterations per step; 0 = sized so that the timed steps last --min-seconds")
    ap.add_argument("--min-seconds", type=float, default=3.5, help="minimum length of the timed region (the clock sampler needs >= 3 s)")
    ap.add_argument("--init-iters", type=int, default=100, help="iterations of the random-init phase timed on their own (roofline.init_phase)")
    ap.add_argument("--burnin", type=int, default=3000, help="untimed iterations of the 4K pair before the warm-up steps (incl. --init-iters)")
    ap.add_argument("--batch-pairs", type=int, default=32, help="640x480 pairs per GPU of the batch640 record (0 = skip it)")
    ap.add_argument("--batch-burnin", type=int, default=3000)
    ap.add_argument("--e2e-its", type=int, default=30000, help="options.its of the end-to-end call (optical_flow.m:17: 30000)")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--e2e-budget", type=float, default=150.0, help="seconds the end-to-end call may take at the measured iteration time")
    ap.add_argument("--ref-iters", type=int, default=1, help="reference arm: whole iterations of the 4K pair per step")
    ap.add_argument("--band-transport", default=None, choices=["p2p", "nccl"], help="halo/sum transport at N>1 (default p2p)")
    ap.add_argument("--band-rows", type=int, default=M_4K)
    ap.add_argument("--band-cols", type=int, default=N_4K)
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
