#!/usr/bin/env python
"""bench.py -- QGMAP pixel-iterations/s on B200 (contract: see the task statement; design notes in DESIGN.md section 5).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA kernels through the C ABI of libqgmap.so)
  python bench.py --impl reference --gpus N --steps K ...   reference arm: the CPU fp64 restatement on the host cores

Workload (BASELINE.json configs[1]): the 8 Middlebury training sequences' shapes (388x584 x3, 380x420, 480x640 x4),
L=2 mixture components, K=9 (9x9 Gauss-Hermite, the reference driver's default optical_flow.m:16), synthetic frames.
One STEP = `--iters` ascent iterations (default 200) on each of the 8 frame pairs.  pixel-iterations/s = pixels x
iterations / seconds.  With N GPUs every rank runs its own 8 pairs (frame pairs are independent: no data-path
collective; weak scaling), whole-job value = N x units / max-over-ranks time.
"""
import argparse
import importlib
import json
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

L_MIX, K_GH = 2, 9
SEQS = ["RubberWhale", "Dimetrodon", "Hydrangea", "Venus", "Grove2", "Grove3", "Urban2", "Urban3"]


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


def workload(pkg, seed0=1234):
    """Synthetic frame pairs of the 8 Middlebury shapes + reference-style options (optical_flow.m:16-23 with L=2)."""
    items = []
    for i, name in enumerate(SEQS):
        M, N = pkg.middlebury_shapes[name]
        I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N, seed=seed0 + 2 * i)
        opts = dict(K=K_GH, L=L_MIX, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0,
                    minu=minu, maxu=maxu, minv=minv, maxv=maxv)
        items.append((name, I1, I2, flow, opts))
    return items


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


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the QGMAP path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module(PKG)
    items = workload(pkg, seed0=1234 + 100 * rank)
    px = sum(I1.size for _, I1, _, _, _ in items)
    iters = args.iters

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg: `value` ----------------
    solvers = []
    for i, (name, I1, I2, flow, opts) in enumerate(items):
        o = dict(opts, device=local)
        s = pkg.Solver(o, I1, I2)
        s.init_state(seed=4321 + i)
        solvers.append(s)
    if args.burnin > 0:                      # untimed: leave the chaotic first phase of the ascent
        pkg.batch_step(solvers, args.burnin)
    for _ in range(args.warmup):
        pkg.batch_step(solvers, iters)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    dev_ms, launches = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ms, nl = pkg.batch_step(solvers, iters)
        dev_ms += ms
        launches += nl
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    # per-kernel view for the roofline: one sequence alone, kernel time from CUDA events on its own stream
    big = max(range(len(items)), key=lambda i: items[i][1].size)
    r = solvers[big].step(iters)
    kern_ms = r["ms"] / max(r["n_done"], 1)
    kern_px = items[big][1].size
    for s in solvers:
        s.close()
    t_dev = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t_dev.item())
    value = world * px * iters * args.steps / (dev_ms_max * 1e-3)

    # ---------------- end-to-end leg: the public call with HOST buffers ----------------
    its_e2e = args.e2e_its
    e2e_in, h2d, d2h = [], 0, 0
    keep = []
    for i, (name, I1, I2, flow, opts) in enumerate(items):
        M, N = I1.shape
        init = init_arrays(opts, M, N, L_MIX, 999 + i)
        pin = {}
        for k, a in list(init.items()) + [("I1", I1), ("I2", I2)]:
            pin[k], t = pinned_like(torch, a)
            keep.append(t)
            h2d += a.size * 8
        o = dict(opts, its=its_e2e, init={k: pin[k] for k in init}, device=local)
        e2e_in.append((o, pin["I1"], pin["I2"]))
        d2h += (4 * M * N * L_MIX + L_MIX + its_e2e) * 8          # mu, sigma, alpha, Energy
    pkg.gqmap_gpu_mixture(*e2e_in[0])                              # warm (context, graph instantiation)
    barrier()
    t0 = time.perf_counter()
    e2e_launches = 0
    for _ in range(args.e2e_steps):
        for o, a, b in e2e_in:
            mu, sigma, alpha, AEPE, Energy, logP = pkg.gqmap_gpu_mixture(o, a, b)
            e2e_launches += pkg.last_solve_stats()[0]
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * px * its_e2e * args.e2e_steps / float(t_e.item())

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
    F = flops_per_px_it(L_MIX, K_GH)
    B = bytes_per_px_it(L_MIX)
    ach_tf = kern_px * F / (kern_ms * 1e-3) / 1e12
    ach_gbs = kern_px * B / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    out = {
        "metric": "QGMAP pixel-iterations/s", "value": value, "unit": "pixel-iter/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "8 Middlebury-shaped synthetic frame pairs (388x584 x3, 380x420, 480x640 x4), L=2 mixture, "
                               "K=9 (9x9 Gauss-Hermite), gqmap_gpu_mixture path; %d iterations per pair per step" % iters,
                   "pixels_per_step": px, "iters_per_step": iters, "L": L_MIX, "K": K_GH,
                   "trajectory_window": "timed iterations %d..%d of each pair's ascent from the reference's random init (untimed burn-in %d + "
                                        "%d warm-up steps); the gather gets more coherent as the beliefs converge, a 30000-iteration "
                                        "solve spends >85%% of its iterations past this window" % (
                                            args.burnin + args.warmup * iters + 1, args.burnin + (args.warmup + args.steps) * iters,
                                            args.burnin, args.warmup),
                   "l2": "no explicit flush: the working set (8 pairs x 2 ping-pong state buffers = %.0f MB) is larger than the 126 MB "
                         "L2 and is cycled every iteration because the 8 pairs advance concurrently on 8 streams; the kernel is "
                         "FP32-pipe bound, state traffic is <1%% of its time" % (px * 9 * L_MIX * 4 * 2 / 1e6),
                   "parallelism": "frame pairs sharded by rank, no collective" if world > 1 else "1 GPU, 8 pairs on 8 streams"},
        "clocks": clocks, "gpu_launches": int(launches),
        "wall_s_timed": wall,
        "e2e": {"value": e2e_value, "unit": "pixel-iter/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "its_per_call": its_e2e, "calls_per_step": len(items), "steps": args.e2e_steps, "gpu_launches": int(e2e_launches),
                "api": "gqmap_gpu_mixture(options,I1,I2) -> qgmap_solve (C ABI), pinned host buffers in, host arrays out, "
                       "monitoring (MAP/logP) at it=1 and every 300 as the reference; each call is a whole solve from the random init with "
                       "options.its = %d (the reference drivers use 30000, optical_flow.m:17)" % its_e2e},
        "roofline": {"bound": "fp32", "achieved": ach_tf, "peak": fp32_meas, "unit": "TFLOP/s", "frac": ach_tf / fp32_meas,
                     "peak_source": "FFMA micro-benchmark measured live on this GPU (qgmap_fp32_peak); nominal 148x128x2x1.965GHz = %.1f" % fp32_nominal,
                     "frac_of_nominal": ach_tf / fp32_nominal,
                     "kernel": "qgmap_iter_kernel<9,false,false> on the largest pair (%d px), %.4f ms per launch (CUDA events on its stream)" % (kern_px, kern_ms),
                     "algorithmic": "%.0f flop and %.0f bytes per pixel-iteration (SURVEY 8d)" % (F, B),
                     "traffic": traffic,
                     "hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}},
    }
    if not args.no_cpu and world == 1:                 # reported at N=1 only (rank 0's host cores)
        out["cpu_baseline"] = cpu_baseline(pkg, items, budget_s=args.cpu_budget, nthreads=os.cpu_count() or 1)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_band4k(args):
    """BASELINE configs[3]: ONE synthetic 3840x2160 frame pair, L=3, K=5, split into row bands over the ranks; per iteration
    the boundary rows travel over NCCL/NVLink and 4L doubles are all-reduced (libqgmap's own communicator).  Strong scaling."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module(PKG)
    M, N, L, K = args.band_rows, args.band_cols, 3, 5
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N, seed=1234)
    opts = dict(K=K, L=L, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv,
                maxv=maxv, device=local)
    if world > 1:
        rb, re = pkg.dist.band_rows(M, rank, world)
        opts.update(row_begin=rb, row_end=re)
    s = pkg.Solver(opts, I1, I2)
    transport = "none"
    if world > 1:
        transport = pkg.dist.connect_band(s, dist, device=torch.device("cuda", local), transport=args.band_transport)
    s.init_state(seed=4321)
    iters = args.iters

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    if args.burnin > 0:
        s.step(args.burnin)
    for _ in range(args.warmup):
        s.step(iters)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ms, launches = 0.0, 0
    for _ in range(args.steps):
        r = s.step(iters)
        ms += r["ms"]
        launches += r["launches"]
    barrier()
    clocks = sampler.stop()
    s.close()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    if rank == 0:
        value = M * N * iters * args.steps / (ms_max * 1e-3)
        F = flops_per_px_it(L, K)
        fp32_meas = pkg.fp32_peak(local)
        ach = value * F / 1e12 / world
        print(json.dumps({
            "metric": "QGMAP pixel-iterations/s", "value": value, "unit": "pixel-iter/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "one synthetic %dx%d frame pair, L=3, K=5, gqmap_gpu_mixture path, %d row band(s), 1-row halo exchange + "
                                   "4L-double global sum per iteration, transport %s (p2p = libqgmap's publish kernel over NVLink peer memory, "
                                   "nccl = send/recv + all-reduce); %d iterations per step after %d burn-in" % (N, M, world, transport, iters, args.burnin),
                       "transport": transport,
                       "iters_per_step": iters, "parallelism": "row bands x%d" % world,
                       "l2": "state 2 x %.0f MB + gather layout %.0f MB exceed the 126 MB L2" % (M * N * 9 * L * 4 / 1e6 / world, (M + 2) * N * 32 / 1e6)},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": ach, "peak": fp32_meas, "unit": "TFLOP/s", "frac": ach / fp32_meas,
                         "note": "per GPU; algorithmic %.0f flop per pixel-iteration" % F}}))
    if world > 1:
        dist.destroy_process_group()


def _gen_pair(seed):
    sys.path.insert(0, os.path.join(ROOT, PKG))
    import frames
    return frames.synthetic_pair(480, 640, seed=seed)


def run_batch256(args):
    """BASELINE configs[4]: a batch of 256 synthetic 640x480 frame pairs (L=3, K=5; pair p uses seed 1234+2p) sharded by
    pair over the ranks, throughput mode, no communication.  Strong scaling: total work fixed, 256/N pairs per GPU."""
    import torch
    import torch.distributed as dist
    from multiprocessing import get_context
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module(PKG)
    npairs = args.batch_pairs
    mine = pkg.dist.shard_pairs(npairs, rank, world)
    with get_context("spawn").Pool(max(1, min(len(mine), (os.cpu_count() or 8) // world))) as pool:      # host-side workload generation
        frames_ = pool.map(_gen_pair, [1234 + 2 * p for p in mine])
    L, K = 3, 5
    solvers = []
    for p_, (I1, I2, flow, (minu, maxu, minv, maxv)) in zip(mine, frames_):
        o = dict(K=K, L=L, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, device=local)
        s = pkg.Solver(o, I1, I2)
        s.init_state(seed=4321 + p_)
        solvers.append(s)
    del frames_
    iters = args.iters

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    if args.burnin > 0:
        pkg.batch_step(solvers, args.burnin)
    for _ in range(args.warmup):
        pkg.batch_step(solvers, iters)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ms, launches = 0.0, 0
    for _ in range(args.steps):
        m_, nl = pkg.batch_step(solvers, iters)
        ms += m_
        launches += nl
    barrier()
    clocks = sampler.stop()
    for s in solvers:
        s.close()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    if rank == 0:
        value = npairs * 480 * 640 * iters * args.steps / (ms_max * 1e-3)
        F = flops_per_px_it(L, K)
        fp32_meas = pkg.fp32_peak(local)
        ach = value * F / 1e12 / world
        print(json.dumps({
            "metric": "QGMAP pixel-iterations/s", "value": value, "unit": "pixel-iter/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "batch of %d synthetic 640x480 frame pairs (seeds 1234+2p), L=3, K=5, gqmap_gpu_mixture path, sharded by pair over "
                                   "%d GPU(s), no communication; %d iterations per pair per step after %d burn-in" % (npairs, world, iters, args.burnin),
                       "pairs_per_gpu": len(mine), "iters_per_step": iters, "parallelism": "pairs sharded by rank, no collective",
                       "l2": "per GPU %d pairs x 2 x 33 MB state buffers cycle through HBM every iteration (>> 126 MB L2)" % len(mine)},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": ach, "peak": fp32_meas, "unit": "TFLOP/s", "frac": ach / fp32_meas,
                         "note": "per GPU; algorithmic %.0f flop per pixel-iteration" % F}}))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(pkg, items, budget_s=12.0, nthreads=0):
    """The oracle (kind 'port': fp64 C restatement of gqmap_gpu_mixture.m, OpenMP) timed on the host cores on a bounded
    sample: the RubberWhale-shaped pair of the same workload, as many whole iterations as fit the budget."""
    from oracle import oracle as O
    name, I1, I2, flow, opts = items[0]
    M, N = I1.shape
    cfg = O.make_config(M, N, opts["L"], opts["K"], lambdas=opts["lambdas"], minu=opts["minu"], maxu=opts["maxu"],
                        minv=opts["minv"], maxv=opts["maxv"], nthreads=nthreads)
    st = O.init_state(cfg, 4321)
    VV = O.get_vv(I2)
    t0 = time.perf_counter()
    O.run(cfg, I1, VV, st, 1, 10 ** 6, 1)
    t1 = time.perf_counter() - t0
    n = max(1, min(50, int(budget_s / max(t1, 1e-3)) - 1))
    t0 = time.perf_counter()
    O.run(cfg, I1, VV, st, 2, 10 ** 6, n)
    dt = time.perf_counter() - t0
    cores = os.cpu_count() if nthreads == 0 else nthreads
    return {"value": M * N * n / dt, "unit": "pixel-iter/s", "cores": cores, "kind": "port",
            "sample": "%d iterations of the %s-shaped pair (%dx%d), L=%d K=%d, fp64 C restatement of gqmap_gpu_mixture.m with OpenMP "
                      "(the MATLAB reference cannot run here)" % (n, name, M, N, opts["L"], opts["K"])}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path = the fp64 oracle port, all host threads, bounded
    sample of the same workload per step (1 iteration on each of the 8 pairs)."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, PKG))
    import frames                                       # workload generation only (no libqgmap, no GPU)
    from oracle import oracle as O

    class _P:
        middlebury_shapes = frames.middlebury_shapes
        synthetic_pair = staticmethod(frames.synthetic_pair)
    items = workload(_P)[:args.ref_pairs]
    probs = []
    for i, (name, I1, I2, flow, opts) in enumerate(items):
        M, N = I1.shape
        cfg = O.make_config(M, N, opts["L"], opts["K"], lambdas=opts["lambdas"], minu=opts["minu"], maxu=opts["maxu"],
                            minv=opts["minv"], maxv=opts["maxv"], nthreads=os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1
        probs.append((cfg, I1, O.get_vv(I2), O.init_state(cfg, 4321 + i)))
    px = sum(p[1].size for p in probs)
    it = 1
    iters = args.ref_iters

    def step():
        nonlocal it
        for cfg, I1, VV, st in probs:
            O.run(cfg, I1, VV, st, it, 10 ** 6, iters)
        it += iters
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = px * iters * args.steps / dt
    cores = os.cpu_count()
    sample = ("%d iteration(s) per step on each of the first %d of the 8 Middlebury-shaped synthetic pairs (%d px), L=%d K=%d, fp64 C restatement "
              "of gqmap_gpu_mixture.m (oracle port; MATLAB/Octave absent), OpenMP on all host threads" % (iters, len(probs), px, L_MIX, K_GH))
    out = {"impl": "reference", "metric": "QGMAP pixel-iterations/s", "value": v, "unit": "pixel-iter/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "8 Middlebury-shaped synthetic frame pairs, L=2, K=9, gqmap_gpu_mixture path (CPU restatement)",
                      "pixels_per_step": px, "iters_per_step": iters},
           "cpu_baseline": {"value": v, "unit": "pixel-iter/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "pixel-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=200, help="ascent iterations per frame pair per step (device-resident leg)")
    ap.add_argument("--burnin", type=int, default=3000, help="untimed iterations per pair before the warm-up steps")
    ap.add_argument("--e2e-its", type=int, default=30000, help="options.its of each end-to-end gqmap_gpu_mixture call (optical_flow.m:17: 30000)")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--ref-iters", type=int, default=1, help="reference arm: iterations per pair per step")
    ap.add_argument("--workload", default="middlebury8", choices=["middlebury8", "band4k", "batch256"],
                    help="middlebury8 = BASELINE configs[1] (default, independent pairs sharded by rank); band4k = configs[3] (one 4K pair in "
                         "row bands); batch256 = configs[4] (256 synthetic 640x480 pairs sharded by pair)")
    ap.add_argument("--batch-pairs", type=int, default=256)
    ap.add_argument("--band-transport", default=None, choices=["p2p", "nccl"], help="band4k: halo/sum transport (default p2p)")
    ap.add_argument("--band-rows", type=int, default=2160)
    ap.add_argument("--band-cols", type=int, default=3840)
    ap.add_argument("--ref-pairs", type=int, default=8, help="reference arm: how many of the 8 pairs form the bounded sample")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "band4k":
        run_band4k(args)
    elif args.workload == "batch256":
        run_batch256(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
