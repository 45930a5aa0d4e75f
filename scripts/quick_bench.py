"""Quick device-time probe of the iteration kernel (development aid; bench.py is the contract)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("gqmap-opticalflow_b200")

def run(M, N, L, K, variant, n=50, warm=10):
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N)
    opts = dict(K=K, L=L, temperature=0.2 if variant == "super" else 0.0, drate=0.75, epsn=1e-6, lambdad=1.0,
                lambdas=16.0 if variant == "super" else 5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, its=10**6)
    with pkg.Solver(opts, I1, I2, variant=variant) as s:
        s.init_state(1)
        s.step(warm)
        r = s.step(n)
        px = M * N
        print("%-5s %4dx%-4d L=%d K=%2d: %8.3f ms/it  %7.3f Gpx-it/s  E=%.6e" % (variant, M, N, L, K, r["ms"] / n, px * n / r["ms"] / 1e6, r["Energy"][-1]), flush=True)

if __name__ == "__main__":
    run(388, 584, 1, 3, "full")
    run(480, 640, 2, 9, "full")
    run(480, 640, 3, 5, "full")
    run(480, 640, 3, 5, "super")
    run(2160, 3840, 3, 5, "full", n=10, warm=3)
