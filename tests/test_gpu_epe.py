"""north_star: "final flow endpoint error within 1e-3 px of the reference's CPU path".  AEPE against the Middlebury ground truth is
the reference's end-to-end signal (gqmap_gpu_mixture.m:63-64); the MAP flow is what get_map_mex extracts (:52-58).

What is measured first (scripts/epe_gate_probe.py, profiles/r02_epe_gate.txt): the ascent never contracts -- at it = 20000 the mean
|G_mu| is still 10..90 and the step 0.03, beliefs move by a pixel per iteration inside the clamp range -- so ANY difference, also one
fp32 rounding in the fp64 oracle itself, grows by about 2x per iteration and saturates after ~20 iterations at the size of the
attractor (0.2..0.5 px mean endpoint distance).  The gate is therefore written where it is well-posed:
  (1) from the SAME late state, after one and two iterations: MAP flows within 1e-3 px in the mean and for >= 99% of the pixels
      (the rest are argmax flips between mixture modes, :54-58), |dAEPE| <= 1e-3 px -- the north_star figure;
  (2) after a longer window: the CUDA flow is as far from the oracle's as the oracle restarted one fp32 rounding away is
      (same mean endpoint distance within 25%), and AEPE agrees to 5e-3 px;
  (3) early phase, statistically: over 8 seeds |AEPE_cuda - AEPE_oracle| stays inside the oracle's own seed-to-seed spread.
Sequences: the committed RubberWhale crop always; the full RubberWhale / Venus / Grove2 pairs when data/_middlebury/ (a git-ignored
copy of the reference's data, scripts/middlebury_pack.py) travelled with the repo."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sequences(pkg):
    d = np.load(os.path.join(ROOT, "tests", "golden", "rubberwhale_crop.npz"))
    out = [("RubberWhale-crop", d["I1"].astype(np.float64), d["I2"].astype(np.float64), d["flow"].astype(np.float64))]
    for name in ("RubberWhale", "Venus", "Grove2"):
        fn = os.path.join(ROOT, "data", "_middlebury", name + ".npz")
        if os.path.exists(fn):
            z = np.load(fn)
            out.append((name, pkg.rgb2gray(z["frame10"]).astype(np.float64), pkg.rgb2gray(z["frame11"]).astype(np.float64),
                        z["flow10"].astype(np.float64)))
    return out


def _setup(pkg, O, I1, I2, raw, L, K):
    I1, I2 = np.asfortranarray(I1), np.asfortranarray(I2)
    img, tflow, minu, maxu, minv, maxv, unk = pkg.flowToColor_mex(np.asfortranarray(raw))      # optical_flow.m:12-13
    M, N = I1.shape
    opts = dict(K=K, L=L, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
    cfg = O.make_config(M, N, L, K, lambdas=5.0, epsn=1e-6, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
    return I1, I2, tflow, unk, opts, cfg, O.get_vv(I2)


def _f32(a):
    return np.asfortranarray(np.asarray(a).astype(np.float32).astype(np.float64))


def _oracle_state(O, g):
    return O.State(_f32(g["muu"]), _f32(g["muv"]), _f32(g["sigmau"]), _f32(g["sigmav"]), _f32(g["pn"]), _f32(g["rou"]), g["w"],
                   alpha=g["alpha"], T=g["T"])


def _dist(a, b):
    return np.sqrt(((a - b) ** 2).sum(axis=2))[1:-1, 1:-1]


def test_late_state_map_flow_and_aepe(pkg, O):
    L, K, burn, window = 2, 5, 20000, 40
    for name, I1, I2, raw in _sequences(pkg):
        I1, I2, tflow, unk, opts, cfg, VV = _setup(pkg, O, I1, I2, raw, L, K)
        with pkg.Solver(opts, I1, I2) as s:
            s.init_state(3)
            s.step(burn)
            late = s.get_state()
            ref, per = _oracle_state(O, late), _oracle_state(O, late)
            per.muu *= (1 + 2.0 ** -24)                            # the oracle restarted ONE fp32 rounding away
            it, done = int(late["it"]), 0
            s.set_state(dict(muu=ref.muu, muv=ref.muv, sigmau=ref.sigu, sigmav=ref.sigv, pn=ref.pn, rou=ref.rou, w=ref.w), T=ref.T, it=it,
                        alpha=ref.alpha)
            for n in (1, 2, window):
                O.run(cfg, I1, VV, ref, it + done, 10 ** 9, n - done)
                O.run(cfg, I1, VV, per, it + done, 10 ** 9, n - done)
                s.step(n - done)
                done = n
                mo = O.find_map(ref.alpha, ref.muu, ref.sigu, ref.muv, ref.sigv)
                mp = O.find_map(per.alpha, per.muu, per.sigu, per.muv, per.sigv)
                mg = s.map()
                ao, ag = O.aepe(cfg, mo, tflow, unk), s.aepe(mg, tflow, unk)
                dg, dp = _dist(mo, mg), _dist(mo, mp)
                info = (name, n, float(dg.mean()), float(np.median(dg)), float((dg > 1e-3).mean()), abs(ag - ao), float(dp.mean()))
                if n <= 2:                                         # (1) the north_star gate
                    assert dg.mean() <= 1e-3 and np.median(dg) <= 1e-4 and (dg > 1e-3).mean() <= 0.01 and abs(ag - ao) <= 1e-3, info
                else:                                              # (2) as close to the oracle as the oracle is to itself
                    assert 0.75 * dp.mean() - 1e-3 <= dg.mean() <= 1.25 * dp.mean() + 1e-3 and abs(ag - ao) <= 5e-3, info


def test_early_phase_aepe_statistics(pkg, O):
    d = np.load(os.path.join(ROOT, "tests", "golden", "rubberwhale_crop.npz"))
    I1, I2, tflow, unk, opts, cfg, VV = _setup(pkg, O, d["I1"][20:68, 30:94].astype(np.float64), d["I2"][20:68, 30:94].astype(np.float64),
                                               d["flow"][20:68, 30:94].astype(np.float64), 2, 3)
    its, A_o, A_g = 400, [], []
    for seed in range(8):
        st = O.init_state(cfg, 100 + seed)
        for f in ("muu", "muv", "sigu", "sigv"):
            getattr(st, f)[...] = _f32(getattr(st, f))
        with pkg.Solver(opts, I1, I2) as s:
            s.set_state(dict(muu=st.muu, muv=st.muv, sigmau=st.sigu, sigmav=st.sigv, pn=st.pn, rou=st.rou, w=st.w), T=0.0)
            s.step(its)
            A_g.append(s.aepe(s.map(), tflow, unk))
        O.run(cfg, I1, VV, st, 1, 10 ** 9, its)
        A_o.append(O.aepe(cfg, O.find_map(st.alpha, st.muu, st.sigu, st.muv, st.sigv), tflow, unk))
    A_o, A_g = np.array(A_o), np.array(A_g)
    spread = A_o.std(ddof=1)
    info = (A_o.round(4).tolist(), A_g.round(4).tolist(), float(spread))
    assert abs(A_g.mean() - A_o.mean()) <= spread, info                      # the two samples of 8 come from the same distribution
    assert np.abs(A_g - A_o).max() <= 4 * spread + 1e-3, info
    assert 0.5 * spread <= A_g.std(ddof=1) <= 2.0 * spread, info
