"""Worker for tests/test_multiprocess_cpu.py (world_size 2, gloo, CPU): validates the row-band PROTOCOL of SURVEY 8e
-- which rows travel, which sums are reduced, in which order -- with the NumPy twin standing in for the CUDA kernel, and the
host-side plumbing (band_rows, shard_pairs, unique-id broadcast, assemble_bands)."""
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
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = importlib.import_module("gqmap-opticalflow_b200")
    from oracle import oracle as O, numpy_twin as T
    from conftest import make_problem

    # ---- plumbing ----
    uid = pkg.dist.broadcast_unique_id(dist)
    ids = [None] * world
    dist.all_gather_object(ids, uid)
    assert len(uid) == 128 and all(i == ids[0] for i in ids) and any(b != 0 for b in uid)
    assert sum(len(pkg.dist.shard_pairs(8, r, world)) for r in range(world)) == 8
    assert pkg.dist.shard_pairs(8, rank, world) == list(range(4 * rank, 4 * rank + 4))

    # ---- band protocol, emulated with the NumPy twin ----
    for sup in (False, True):
        Mo, No, L, K, Tm = ((26, 22) if not sup else (48, 40)) + (2, 3, 0.2)
        cfg, I1, I2, st = make_problem(O, Mo, No, L, K, super=sup, seed=17, T=Tm, small_sigma=True)
        VV = np.asarray(O.get_vv(I2)); I1 = np.asarray(I1)
        M, N = cfg.M, cfg.N
        rb, re = pkg.dist.band_rows(M, rank, world)
        g0, g1 = max(rb - 1, 0), min(re + 1, M)                      # stored rows: owned + one halo row each side
        loc = {k: np.array(getattr(st, k)[g0:g1]) for k in ("muu", "muv", "sigu", "sigv", "pn", "rou")}
        alpha, w = st.alpha.copy(), st.w.copy()
        ref = st.copy()
        nit, it = 4, 499                                             # crosses it=500 -> alpha update active
        _, _, _, Eref, dmref, _ = O.run(cfg, I1, np.asfortranarray(VV), ref, it, 10 ** 6, nit)
        lo, hi = max(rb, 1) - g0, min(re, M - 1) - g0               # local indices of the rows this rank updates
        for k in range(nit):
            step = cfg.step0 / (1 + (it + k) / cfg.step_tau)
            d = dict(loc, alpha=alpha)
            t = T.iteration_gradients(I1, VV, d, K=K, T=Tm, lambdad=cfg.lambdad, lambdas=cfg.lambdas, epsn=cfg.epsn,
                                      super_=sup, guard_a0=not sup, row_off=g0)
            rows = (slice(lo, hi), slice(1, N - 1))
            # local partial sums of :36,:48,:69 over the rows this rank owns
            e_px = t["nEnergy"] + t["eEnergy"].sum(axis=(3, 4)); da_px = t["dan"] + t["dae"].sum(axis=(3, 4))
            sums = np.stack([e_px[rows].sum(axis=(0, 1)), da_px[rows].sum(axis=(0, 1)),
                             np.abs(t["G_muu"][rows]).sum(axis=(0, 1)), np.abs(t["G_sigu"][rows]).sum(axis=(0, 1))])
            ts = torch.from_numpy(sums.copy())
            dist.all_reduce(ts)                                      # the 4L-double all-reduce
            tot = ts.numpy()
            cl = lambda v, a, b: np.minimum(np.maximum(v, a), b)
            new = {k2: v.copy() for k2, v in loc.items()}
            new["muu"][rows] = cl(loc["muu"][rows] + t["G_muu"][rows] * step, cfg.minu, cfg.maxu)
            new["muv"][rows] = cl(loc["muv"][rows] + t["G_muv"][rows] * step, cfg.minv, cfg.maxv)
            new["sigu"][rows] = cl(loc["sigu"][rows] + t["G_sigu"][rows] * step, 0.01, cfg.sigma_max)
            new["sigv"][rows] = cl(loc["sigv"][rows] + t["G_sigv"][rows] * step, 0.01, cfg.sigma_max)
            new["pn"][rows] = cl(loc["pn"][rows] + t["dpn"][rows] * step, -cfg.corr_tor, cfg.corr_tor)
            new["rou"][rows] = cl(loc["rou"][rows] + t["drou"][rows] * step, -cfg.corr_tor, cfg.corr_tor)
            loc = new
            E = tot[0].sum()
            assert abs(E / Eref[k] - 1) < 1e-12, (rank, k, E, Eref[k])
            assert abs(tot[2].sum() / ((M - 2) * (N - 2) * L) / dmref[k] - 1) < 1e-12
            if it + k > cfg.alpha_start:                             # every rank applies the same alpha update (:78-86)
                dal = tot[1]
                dw = alpha * (dal - (dal * alpha).sum())
                w = np.clip(w + dw * step * cfg.alpha_scale, -300, 300)
                alpha = np.exp(w) / np.exp(w).sum()
            if cfg.anneal_every > 0 and (it + k) % cfg.anneal_every == 0:       # super-pixel variant, S:72
                Tm = max(Tm * cfg.drate, cfg.T_floor)
            # halo exchange: whole boundary rows of all fields, up and down
            reqs, bufs = [], {}
            if rank > 0:
                send = np.concatenate([loc[f][rb - g0].ravel() for f in sorted(loc)])
                bufs["up"] = torch.zeros(send.size, dtype=torch.float64)
                reqs += [dist.isend(torch.from_numpy(send.copy()), rank - 1), dist.irecv(bufs["up"], rank - 1)]
            if rank < world - 1:
                send = np.concatenate([loc[f][re - 1 - g0].ravel() for f in sorted(loc)])
                bufs["dn"] = torch.zeros(send.size, dtype=torch.float64)
                reqs += [dist.isend(torch.from_numpy(send.copy()), rank + 1), dist.irecv(bufs["dn"], rank + 1)]
            for r in reqs:
                r.wait()
            for key, row in (("up", rb - 1 - g0), ("dn", re - g0)):
                if key in bufs:
                    off = 0
                    for f in sorted(loc):
                        n_el = loc[f][row].size
                        loc[f][row] = bufs[key].numpy()[off:off + n_el].reshape(loc[f][row].shape)
                        off += n_el
        full = pkg.dist.assemble_bands(dist, {k: _pad(v, g0, M) for k, v in loc.items()}, M)
        for f in ("muu", "muv", "sigu", "sigv", "pn", "rou"):
            dd = np.abs(full[f] - getattr(ref, f)); assert dd.max() < 1e-8, (rank, f, dd.max(), np.unravel_index(dd.argmax(), dd.shape), rb, re)
        assert np.abs(alpha - ref.alpha).max() < 1e-15
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


def _pad(v, g0, M):
    out = np.zeros((M,) + v.shape[1:])
    out[g0:g0 + v.shape[0]] = v
    return out


if __name__ == "__main__":
    main()
