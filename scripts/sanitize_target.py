"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck): both variants, both lane mappings,
band group, MAP/logP/AEPE, solve."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("gqmap-opticalflow_b200")
for lanes in ("1", "4"):
    os.environ["QGMAP_LANES"] = lanes
    for variant, (M, N), L, K in (("full", (37, 45), 2, 3), ("full", (20, 70), 3, 4), ("super", (48, 64), 2, 5)):
        I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N)
        opts = dict(K=K, L=L, temperature=0.2, drate=0.75, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
        with pkg.Solver(opts, I1, I2, variant=variant) as s:
            s.init_state(1)
            r = s.step(30)
            m = s.map(); s.logp(m); s.aepe(m, flow, np.zeros((M, N), bool)); s.debug_gradients(); s.get_state()
        with pkg.BandGroup(opts, I1, I2, 3, variant=variant) as g:
            g.init_state(1)
            rb = g.step(30)
        assert np.abs(rb["Energy"] / r["Energy"] - 1).max() < 1e-10
        fn = pkg.gqmap_gpuSuper_mix_entropy if variant == "super" else pkg.gqmap_gpu_mixture
        fn(dict(opts, its=12, seed=2, trueFlow=flow, unknownIdx=np.zeros((M, N), bool), log_every=5), I1, I2)
        print("ok", lanes, variant, M, N, L, K, flush=True)
