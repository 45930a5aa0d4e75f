"""Development probe: the peer-memory band transport forced onto ONE GPU (QGMAP_GROUP_TRANSPORT=p2p) against the single domain."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
os.environ["QGMAP_GROUP_TRANSPORT"] = sys.argv[1] if len(sys.argv) > 1 else "p2p-shared"
pkg = importlib.import_module("gqmap-opticalflow_b200")
from oracle import oracle as O
from conftest import make_problem, options_from_cfg, state_dict
for variant, (Mo, No), L, K, T, nb in (("full", (61, 70), 2, 3, 0.0, 2), ("full", (96, 45), 3, 5, 0.2, 3), ("super", (128, 96), 3, 5, 0.2, 2)):
    cfg, I1, I2, st = make_problem(O, Mo, No, L, K, super=variant == "super", seed=5, T=T, small_sigma=True)
    opts = options_from_cfg(cfg, T=T, alpha_scale=1e-5)
    with pkg.Solver(opts, I1, I2, variant=variant) as s:
        s.set_state(state_dict(st), T=T, it=495)
        r1 = s.step(14)
        a = s.get_state()
    t0 = time.time()
    with pkg.BandGroup(opts, I1, I2, nb, variant=variant) as g:
        g.set_state(state_dict(st), T=T, it=495)
        ra = g.step(6)
        rb = g.step(8)
        b = g.get_state()
    ok = all(np.array_equal(a[f], b[f]) for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou"))
    E = np.concatenate([ra["Energy"], rb["Energy"]])
    print(variant, nb, "bands: identical", ok, "energy rel", np.abs(E / r1["Energy"] - 1).max(), "alpha", np.abs(a["alpha"] - b["alpha"]).max(),
          "n_done", ra["n_done"], rb["n_done"], "%.2fs" % (time.time() - t0), flush=True)
