"""ms/iteration along a long trajectory (how the gather coherence evolves).  usage: traj_bench.py [variant L K M N total window]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("gqmap-opticalflow_b200")
a = sys.argv[1:]
variant = a[0] if len(a) > 0 else "full"
L = int(a[1]) if len(a) > 1 else 2
K = int(a[2]) if len(a) > 2 else 9
M = int(a[3]) if len(a) > 3 else 480
N = int(a[4]) if len(a) > 4 else 640
total = int(a[5]) if len(a) > 5 else 10000
win = int(a[6]) if len(a) > 6 else 500
I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N)
opts = dict(K=K, L=L, temperature=0.2 if variant == "super" else 0.0, drate=0.75, epsn=1e-6, lambdad=1.0,
            lambdas=16.0 if variant == "super" else 5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
unk = np.zeros((M, N), bool)
with pkg.Solver(opts, I1, I2, variant=variant) as s:
    s.init_state(1)
    it = 0
    while it < total:
        r = s.step(win)
        it += r["n_done"]
        st = s.get_state()
        m = s.map()
        sc = 4 if variant == "super" else 1
        aepe = s.aepe(m, flow, unk)
        print("it %6d  %.4f ms/it  %.3f Gpx-it/s  E=%.5e  ptdmu=%.3e  mean sig_u=%.3f  AEPE=%.4f  alpha=%s" % (
            it, r["ms"] / r["n_done"], M * N * r["n_done"] / r["ms"] / 1e6, r["Energy"][-1], r["ptdmu"][-1],
            st["sigmau"].mean(), aepe, np.round(st["alpha"], 3)), flush=True)
        if r["stopped"]:
            break
