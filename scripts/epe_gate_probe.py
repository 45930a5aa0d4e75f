"""Where the north_star EPE gate (final-flow endpoint error within 1e-3 px of the reference's CPU path) is well-posed: the LATE phase.
A late GPU state (it = burn) of a Middlebury sequence is handed to the fp64 oracle and to the CUDA path; both run n more iterations;
their MAP flows (get_map_mex) and AEPE against the ground truth (gqmap_gpu_mixture.m:63-64) are compared.
usage: epe_gate_probe.py [burn] [n] [names,...] [L] [K]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as O
pkg = importlib.import_module("gqmap-opticalflow_b200")
burn = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
names = (sys.argv[3] if len(sys.argv) > 3 else "RubberWhale,Venus,Grove2").split(",")
L = int(sys.argv[4]) if len(sys.argv) > 4 else 2
K = int(sys.argv[5]) if len(sys.argv) > 5 else 5
for name in names:
    d = np.load(os.path.join(ROOT, "data", "_middlebury", name + ".npz"))
    I1 = np.asfortranarray(pkg.rgb2gray(d["frame10"]).astype(np.float64))
    I2 = np.asfortranarray(pkg.rgb2gray(d["frame11"]).astype(np.float64))
    img, tflow, minu, maxu, minv, maxv, unk = pkg.flowToColor_mex(np.asfortranarray(d["flow10"].astype(np.float64)))
    M, N = I1.shape
    opts = dict(K=K, L=L, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
    cfg = O.make_config(M, N, L, K, lambdas=5.0, epsn=1e-6, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
    VV = O.get_vv(I2)
    with pkg.Solver(opts, I1, I2) as s:
        s.init_state(3)
        r = s.step(burn)
        late = s.get_state()
    f32 = lambda a: np.asfortranarray(a.astype(np.float32).astype(np.float64))
    mk = lambda g: O.State(f32(g["muu"]), f32(g["muv"]), f32(g["sigmau"]), f32(g["sigmav"]), f32(g["pn"]), f32(g["rou"]), g["w"], alpha=g["alpha"], T=g["T"])
    ref = mk(late)
    per = mk(late)
    per.muu *= (1 + 2.0 ** -24)                          # the oracle's own sensitivity: ONE fp32 rounding in mu_u
    print("%-12s %dx%d L=%d K=%d late it=%d (burn n_done=%d stopped=%s, last mean|G_mu| %.2e)" % (name, M, N, L, K, late["it"], r["n_done"], r["stopped"], r["ptdmu"][-1]), flush=True)
    it = late["it"]
    done = 0
    with pkg.Solver(opts, I1, I2) as s:
        s.set_state(dict(muu=ref.muu, muv=ref.muv, sigmau=ref.sigu, sigmav=ref.sigv, pn=ref.pn, rou=ref.rou, w=ref.w), T=ref.T, it=it, alpha=ref.alpha)
        for n_ in (1, 2, 5, 10, 20, n):
            k = n_ - done
            O.run(cfg, I1, VV, ref, it + done, 10 ** 9, k)
            O.run(cfg, I1, VV, per, it + done, 10 ** 9, k)
            s.step(k)
            done = n_
            mo = O.find_map(ref.alpha, ref.muu, ref.sigu, ref.muv, ref.sigv)
            mp = O.find_map(per.alpha, per.muu, per.sigu, per.muv, per.sigv)
            mg = s.map()
            ao, ap, ag = O.aepe(cfg, mo, tflow, unk), O.aepe(cfg, mp, tflow, unk), s.aepe(mg, tflow, unk)
            dg = np.sqrt(((mo - mg) ** 2).sum(axis=2))[1:-1, 1:-1]
            dp = np.sqrt(((mo - mp) ** 2).sum(axis=2))[1:-1, 1:-1]
            print("   +%3d its: AEPE oracle %.6f | CUDA dAEPE %.2e dist mean %.2e p50 %.2e p99 %.2e max %.2e frac>1e-3 %.3f | perturbed oracle dAEPE %.2e dist mean %.2e p50 %.2e p99 %.2e max %.2e frac>1e-3 %.3f" % (
                n_, ao, abs(ag - ao), dg.mean(), np.median(dg), np.percentile(dg, 99), dg.max(), (dg > 1e-3).mean(),
                abs(ap - ao), dp.mean(), np.median(dp), np.percentile(dp, 99), dp.max(), (dp > 1e-3).mean()), flush=True)
