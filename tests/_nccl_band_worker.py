"""Worker for tests/test_gpu_multirank.py: one process per GPU, row bands of ONE frame pair, both transports: libqgmap's own
peer-memory publish kernel (qgmap_band_p2p_connect, CUDA IPC) and NCCL (qgmap_band_connect).
Rank 0 also solves the undivided problem and checks the assembled band result is bit-identical."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                       # rendezvous only; the data path is libqgmap's own NCCL communicator
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = importlib.import_module("gqmap-opticalflow_b200")
    from oracle import oracle as O
    from conftest import make_problem, options_from_cfg, state_dict
    for transport, variant, (Mo, No), L, K, T in (("p2p", "full", (75, 90), 2, 5, 0.0), ("p2p", "super", (128, 160), 3, 3, 0.2),
                                                  ("nccl", "full", (75, 90), 2, 5, 0.0), ("nccl", "super", (128, 160), 3, 3, 0.2)):
        sup = variant == "super"
        cfg, I1, I2, st = make_problem(O, Mo, No, L, K, super=sup, seed=23, T=T, small_sigma=True)
        rb, re = pkg.dist.band_rows(cfg.M, rank, world)
        opts = options_from_cfg(cfg, T=T, alpha_scale=1e-5, device=local)
        n = 12
        with pkg.Solver(dict(opts, row_begin=rb, row_end=re), I1, I2, variant=variant) as s:
            pkg.dist.connect_band(s, dist, transport=transport)
            s.set_state(state_dict(st), T=T, it=495)
            r = s.step(5)                                   # two step calls: the p2p flags carry a per-call generation
            r2 = s.step(n - 5)
            r = dict(n_done=r["n_done"] + r2["n_done"], Energy=np.concatenate([r["Energy"], r2["Energy"]]), ms=r["ms"] + r2["ms"])
            full = pkg.dist.assemble_bands(dist, s.get_state(), cfg.M)
        assert r["n_done"] == n, r
        if rank == 0:
            with pkg.Solver(opts, I1, I2, variant=variant) as s1:
                s1.set_state(state_dict(st), T=T, it=495)
                r1 = s1.step(n)
                a = s1.get_state()
            for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou"):
                assert np.array_equal(a[f], full[f]), (variant, f, np.abs(a[f] - full[f]).max())
            assert np.abs(r["Energy"] / r1["Energy"] - 1).max() < 1e-12
            assert np.abs(full["alpha"] - a["alpha"]).max() < 1e-14
            print("%s bands %s world=%d: bit-identical to the single domain, %.3f ms/it" % (transport, variant, world, r["ms"] / n), flush=True)
    # the reference-facing call, one rank per band: dist.gqmap_gpu_mixture_bands vs gqmap_gpu_mixture on rank 0 (40 iterations: the
    # fused exchange also runs from CUDA-graph launches; monitoring shares at it = 1, 10, 20, 30, 40)
    Mo, No = 90, 120
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(Mo, No)
    unk = np.zeros((Mo, No), bool)
    unk[40:50, 10:60] = True
    o2 = dict(K=5, L=3, its=40, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv,
              seed=5, log_every=10, trueFlow=flow, unknownIdx=unk, alpha_start=2, alpha_scale=1e-5, device=local)
    b = pkg.dist.gqmap_gpu_mixture_bands(o2, I1, I2, dist)
    rb, re = pkg.dist.band_rows(Mo, rank, world)
    if rank == 0:
        a = pkg.gqmap_gpu_mixture(o2, I1, I2)
        assert np.array_equal(a[0][rb:re], b[0][rb:re]) and np.array_equal(a[1][rb:re], b[1][rb:re]), "own rows of mu / sigma"
        assert np.abs(a[2] - b[2]).max() < 1e-14
        for k in (3, 4, 5):
            m = ~np.isnan(a[k])
            assert np.array_equal(np.isnan(a[k]), np.isnan(b[k])) and np.abs(b[k][m] / a[k][m] - 1).max() < 1e-11, k
        assert m.sum() == 5
        print("gqmap_gpu_mixture_bands world=%d: bit-identical to gqmap_gpu_mixture" % world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
