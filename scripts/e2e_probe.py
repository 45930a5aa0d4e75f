"""Where does the end-to-end time of one gqmap_gpu_mixture call go? (development aid)"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("gqmap-opticalflow_b200")
M, N, L, K = 480, 640, 2, 9
I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N)
opts = dict(K=K, L=L, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, seed=3)
pkg.gqmap_gpu_mixture(dict(opts, its=1), I1, I2)
for its in (1, 2, 299, 300, 301, 600, 2000):
    t0 = time.perf_counter()
    out = pkg.gqmap_gpu_mixture(dict(opts, its=its), I1, I2)
    dt = time.perf_counter() - t0
    nl, ms = pkg.last_solve_stats()
    print("its=%5d  wall %8.2f ms   iteration kernels %8.2f ms   other %7.2f ms   launches %d" % (its, dt * 1e3, ms, dt * 1e3 - ms, nl), flush=True)
with pkg.Solver(opts, I1, I2) as s:
    s.init_state(1)
    s.step(600)
    for name, fn in (("map", lambda: s.map()), ("get_state", lambda: s.get_state()), ("logp", lambda: s.logp(m)), ):
        if name == "logp":
            m = s.map()
        t0 = time.perf_counter(); fn(); print("%-10s %.2f ms" % (name, (time.perf_counter() - t0) * 1e3))
    t0 = time.perf_counter(); s.init_state(2); print("init_state %.2f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter()
with pkg.Solver(opts, I1, I2) as s:
    print("create     %.2f ms" % ((time.perf_counter() - t0) * 1e3))
