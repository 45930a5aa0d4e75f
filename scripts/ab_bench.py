"""A/B device-time probe at a converged state: ms/iteration of several configs after a burn-in (development aid).
usage: ab_bench.py [tag] [short]   (library chosen by QGMAP_LIB_PATH)"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gqmap-opticalflow_b200")
tag = sys.argv[1] if len(sys.argv) > 1 else ""
CASES = [("full", 480, 640, 2, 9, 4000), ("full", 480, 640, 3, 5, 2000), ("full", 388, 584, 1, 3, 1000), ("full", 480, 640, 2, 9, 0),
         ("full", 480, 640, 3, 5, 0), ("super", 480, 640, 3, 5, 1500), ("full", 2160, 3840, 3, 5, 1000), ("full", 480, 640, 3, 7, 2000),
         ("full", 480, 640, 2, 11, 2000), ("full", 480, 640, 2, 6, 2000)]
if len(sys.argv) > 2:
    CASES = CASES[:5]
for variant, M, N, L, K, burn in CASES:
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N)
    opts = dict(K=K, L=L, temperature=0.2 if variant == "super" else 0.0, drate=0.75, epsn=1e-6, lambdad=1.0,
                lambdas=16.0 if variant == "super" else 5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, its=10**6)
    with pkg.Solver(opts, I1, I2, variant=variant) as s:
        s.init_state(1)
        if burn:
            s.step(burn)
        s.step(20)
        n = 200 if M < 1000 else 40
        best = 1e9
        for _ in range(3):
            r = s.step(n)
            best = min(best, r["ms"] / n)
        print("%-6s %-5s %4dx%-4d L=%d K=%2d burn=%4d: %8.4f ms/it  %7.3f Gpx-it/s  E=%.6e" % (
            tag, variant, M, N, L, K, burn, best, M * N / best / 1e6, r["Energy"][-1]), flush=True)
