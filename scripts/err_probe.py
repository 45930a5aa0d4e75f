"""Where does the fp32 gradient error come from?  (development aid)"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import make_problem, options_from_cfg, state_dict
from oracle import oracle as O
pkg = importlib.import_module("gqmap-opticalflow_b200")

def f32(a): return a.astype(np.float32).astype(np.float64)
for (L, K, small, rho) in ((2, 3, True, 0.6), (2, 3, True, 0.0), (2, 5, False, 0.0), (2, 9, True, 0.9)):
    cfg, I1, I2, st = make_problem(O, 48, 64, L, K, seed=13, small_sigma=small, rho=rho)
    for f in ("muu", "muv", "sigu", "sigv", "pn", "rou"):
        getattr(st, f)[...] = f32(getattr(st, f))
    VV = O.get_vv(I2)
    g = O.gradients(cfg, I1, VV, st, assemble=True)
    with pkg.Solver(options_from_cfg(cfg), I1, I2) as s:
        s.set_state(state_dict(st))
        d = s.debug_gradients()
    print("L=%d K=%d small=%s rho=%.1f" % (L, K, small, rho))
    for name, ref in (("G_muu", g["dmuu"]), ("G_sigu", g["dsigmau"]), ("dpn", g["dpn"]), ("drou", g["drou"])):
        a, b = d[name][1:-1, 1:-1], ref[1:-1, 1:-1]
        err = np.abs(a - b)
        i = np.unravel_index(err.argmax(), err.shape)
        m, n, l = i[0] + 1, i[1] + 1, i[2]
        print("  %-6s max|ref| %9.3e  max err %9.3e  med err %9.3e | at %s ref %+.5e got %+.5e  sigu %.4f sigv %.4f pn %+.4f muu %+.3f muv %+.3f"
              % (name, np.abs(b).max(), err.max(), np.median(err), i, b[i], a[i], st.sigu[m, n, l], st.sigv[m, n, l], st.pn[m, n, l], st.muu[m, n, l], st.muv[m, n, l]))
        if name == "drou":
            print("         rou there %+.5f" % st.rou[(m, n) + tuple(i[2:])])
