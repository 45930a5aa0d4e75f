"""Which thread mapping of the super-pixel kernel wins where: QGMAP_LANES=1 vs 4 on a small (480x640) and a large (2160x3840) frame."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gqmap-opticalflow_b200")
for (M, N, burn) in ((480, 640, 1500), (1080, 1920, 600), (2160, 3840, 300)):
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N)
    opts = dict(K=5, L=3, temperature=0.2, drate=0.75, epsn=1e-6, lambdad=1.0, lambdas=16.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, its=10**6)
    for lanes in ("1", "4"):
        os.environ["QGMAP_LANES"] = lanes
        with pkg.Solver(opts, I1, I2, variant="super") as s:
            s.init_state(1)
            s.step(burn)
            best = min(s.step(50)["ms"] / 50 for _ in range(3))
        print("super %4dx%-4d L=3 K=5 lanes=%s after %d its: %.4f ms/it  %.3f Gpx-it/s" % (M, N, lanes, burn, best, M * N / best / 1e6), flush=True)
