"""Short, graph-free run of the iteration kernel for ncu (one GPU): 480x640 pair, L=2 K=9 by default.
usage: profile_target.py [variant L K M N iters warm g]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gqmap-opticalflow_b200")
a = sys.argv[1:]
variant = a[0] if len(a) > 0 else "full"
L = int(a[1]) if len(a) > 1 else 2
K = int(a[2]) if len(a) > 2 else 9
M = int(a[3]) if len(a) > 3 else 480
N = int(a[4]) if len(a) > 4 else 640
iters = int(a[5]) if len(a) > 5 else 12
warm = int(a[6]) if len(a) > 6 else 0
grey = len(a) > 7 and a[7] == "g"          # frames rounded to integer grey levels (fp16 one-sector gather layout)
I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N, grey_levels=grey)
opts = dict(K=K, L=L, temperature=0.2 if variant == "super" else 0.0, drate=0.75, epsn=1e-6, lambdad=1.0,
            lambdas=16.0 if variant == "super" else 5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
with pkg.Solver(opts, I1, I2, variant=variant) as s:
    s.init_state(1)
    if warm:
        s.step(warm)     # move past the initial large-sigma phase (graph launches: not profiled with -k filters + -s)
    tot = 0.0
    for _ in range(iters):
        r = s.step(1)
        tot += r["ms"]
    print("%s %dx%d L=%d K=%d: %.4f ms/iteration over %d single launches, E=%.6e" % (variant, M, N, L, K, tot / iters, iters, r["Energy"][-1]))
