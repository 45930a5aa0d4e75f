"""How fast do two fp64 evaluations of the SAME program drift apart?  The reference's own gqmap_gpu_mixture.m (executed by
oracle/mlab/minimat.py) against the oracle's C restatement, same draws, free-running.  After iteration 1 they differ by fp64
rounding (different libm / summation order); the ascent then amplifies that (development aid, build container only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import make_refsrc_golden as G
from oracle import oracle as O

O.build()
name, its = (sys.argv[1] if len(sys.argv) > 1 else "full_L2K3"), int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = G.run_case(name, its=its, probes=tuple(range(1, its + 1)))
solver, Mo, No, L, K, T, drate, lambdas, _, _, seed = G.CASES[name]
rg = G.CASE_RANGE.get(name, G.RANGE)
sup = solver.endswith("entropy")
cfg = O.make_config(Mo, No, L, K, super=sup, lambdas=lambdas, drate=drate, **rg)
shp = (cfg.M, cfg.N, cfg.L)
w, ru, rv, su, sv = (out["draw%d" % i] for i in range(5))
st = O.State(rg["minu"] + ru.reshape(shp, order="F") * (rg["maxu"] - rg["minu"]), rg["minv"] + rv.reshape(shp, order="F") * (rg["maxv"] - rg["minv"]),
             su.reshape(shp, order="F") + (rg["maxu"] - rg["minu"]), sv.reshape(shp, order="F") + (rg["maxv"] - rg["minv"]),
             np.zeros(shp), np.zeros(shp + (2, 2)), np.ravel(w), T=T)
VV = O.get_vv(out["I2"])
print("# %s: %dx%d L=%d K=%d T=%g, executed reference source vs oracle restatement, free-running from the same draws" % (name, Mo, No, L, K, T))
print("# it   |E_src/E_oracle - 1|   max|mu_u diff| (px)   max|sigma_u diff|   max|rho_edge diff|   max|rho_edge|")
for k in range(1, int(out["it_end"])):
    n, it, stopped, E, dm, ds = O.run(cfg, out["I1"], VV, st, k, 10 ** 6, 1)
    g = lambda f: out["p%d_%s" % (k, f)].reshape(getattr(st, {"sigmau": "sigu"}.get(f, f)).shape, order="F")
    print("%4d   %.2e             %.2e              %.2e            %.2e            %.6f" % (
        k, abs(out["p%d_Energy" % k] / E[0] - 1), np.abs(g("muu") - st.muu).max(), np.abs(g("sigmau") - st.sigu).max(),
        np.abs(g("rou") - st.rou).max(), np.abs(st.rou).max()))
