"""Lean A/B device-time probe (round 2, late): ms/iteration of one configuration for the library selected by QGMAP_LIB_PATH, with the
synthetic frame pair cached in build/ (generating the 4K pair costs ~20 s of host time per process otherwise).
usage: ab3.py <tag> [M N L K burn [n]]"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gqmap-opticalflow_b200")
tag = sys.argv[1]
M, N, L, K, burn = (int(x) for x in (sys.argv[2:7] if len(sys.argv) >= 7 else (2160, 3840, 3, 5, 300)))
n = int(sys.argv[7]) if len(sys.argv) > 7 else (40 if M > 1000 else 200)
cache = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "frames_%dx%d_grey.npz" % (M, N))
if os.path.exists(cache):
    z = np.load(cache)
    I1, I2 = np.asfortranarray(z["I1"].astype(np.float64)), np.asfortranarray(z["I2"].astype(np.float64))
    minu, maxu, minv, maxv = (float(v) for v in z["ext"])
else:
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N, grey_levels=True)
    os.makedirs(os.path.dirname(cache), exist_ok=True)
    np.savez_compressed(cache, I1=I1.astype(np.uint8), I2=I2.astype(np.uint8), ext=np.array([minu, maxu, minv, maxv]))
opts = dict(K=K, L=L, temperature=0.0, drate=0.75, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, its=10 ** 6)
with pkg.Solver(opts, I1, I2) as s:
    s.init_state(1)
    if burn:
        s.step(burn)
    s.step(20)
    best = 1e9
    for _ in range(3):
        r = s.step(n)
        best = min(best, r["ms"] / n)
    print("%-8s full %4dx%-4d L=%d K=%2d burn=%4d grey: %8.4f ms/it  %7.3f Gpx-it/s  E=%.9e" % (
        tag, M, N, L, K, burn, best, M * N / best / 1e6, r["Energy"][-1]), flush=True)
