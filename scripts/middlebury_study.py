"""Run the reference's two experiment drivers on the Middlebury sequences it ships, through the package's public API
(gqmap_gpu_mixture / gqmap_gpuSuper_mix_entropy exactly as optical_flow.m:12-27 / optical_flowSuper.m:15-34 call them):
flowToColor_mex(readFlowFile(...)) for trueFlow / clamp range / unknown mask, rgb2gray frames, the drivers' options.
Needs data/_middlebury/*.npz (scripts/middlebury_pack.py; git-ignored copy of the reference's data).
usage: middlebury_study.py [its] [which]   which: all | full | super | c1 (BASELINE configs[1]) | c2 (configs[2])"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("gqmap-opticalflow_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
its = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
which = sys.argv[2] if len(sys.argv) > 2 else "all"
GT8 = ["RubberWhale", "Dimetrodon", "Hydrangea", "Venus", "Grove2", "Grove3", "Urban2", "Urban3"]


def load(name):
    d = np.load(os.path.join(ROOT, "data", "_middlebury", name + ".npz"))
    I1 = pkg.rgb2gray(d["frame10"]).astype(np.float64)          # optical_flow.m:8-11 (scale = 1: imresize is the identity)
    I2 = pkg.rgb2gray(d["frame11"]).astype(np.float64)
    flow = np.asfortranarray(d["flow10"].astype(np.float64))
    img, tflow, minu, maxu, minv, maxv, unk = pkg.flowToColor_mex(flow)      # optical_flow.m:12-13
    return I1, I2, tflow, unk, (minu, maxu, minv, maxv)


def run(name, variant, K, L, lambdas, T, drate, label):
    I1, I2, tflow, unk, (minu, maxu, minv, maxv) = load(name)
    opts = dict(K=K, L=L, its=its, epsn=0.001 ** 2, lambdas=lambdas, lambdad=1.0, temperature=T, drate=drate, trueFlow=tflow,
                unknownIdx=unk, minu=minu, maxu=maxu, minv=minv, maxv=maxv, seed=1)
    fn = pkg.gqmap_gpuSuper_mix_entropy if variant == "super" else pkg.gqmap_gpu_mixture
    t0 = time.time()
    mu, sigma, alpha, AEPE, Energy, logP = fn(opts, I1, I2)
    dt = time.time() - t0
    nl, ms = pkg.last_solve_stats()
    logged = np.flatnonzero(~np.isnan(AEPE[:, 0]))
    done = int(np.flatnonzero(Energy[:, 0] != 0)[-1]) + 1 if np.any(Energy[:, 0] != 0) else 0
    a = AEPE[logged, 0]
    pick = [i for i in (0, 1, 3, 10, 33, 66, len(a) - 1) if i < len(a)]
    traj = " ".join("%d:%.3f" % (logged[i] + 1, a[i]) for i in sorted(set(pick)))
    print("%-10s %-5s %-22s %4dx%-4d u[%.2f,%.2f] v[%.2f,%.2f] its %5d  AEPE first %.3f best %.3f last %.3f  E_last %.4e logP_last %.4e  "
          "alpha %s  %.2fs wall, %.2fs kernels, %.3f Gpx-it/s e2e | AEPE@it %s" % (
              name, variant, label, I1.shape[0], I1.shape[1], minu, maxu, minv, maxv, done, a[0], np.nanmin(a), a[-1], Energy[done - 1, 0],
              logP[logged[-1], 0], np.round(alpha.ravel(), 3), dt, ms / 1e3, I1.size * done / dt / 1e9, traj), flush=True)


if which in ("all", "c1"):
    for n in GT8:                 # BASELINE configs[1]: all 8 GT sequences, L=2 mixture, gqmap_gpu_mixture path (driver K=9)
        run(n, "full", 9, 2, 5.0, 0.0, 0.5, "c1 L=2 K=9")
if which in ("all", "full"):
    for n in ("Teddy", "Cones"):  # optical_flow.m:3,16-23 as shipped
        run(n, "full", 9, 3, 5.0, 0.0, 0.5, "driver L=3 K=9")
if which in ("all", "c2"):
    for n in ("Urban2", "Grove3"):   # BASELINE configs[2]
        run(n, "super", 5, 3, 16.0, 0.2, 0.75, "c2 L=3 K=5")
if which in ("all", "super"):
    for n in ("Venus", "Hydrangea", "Urban2", "Urban3", "Grove3"):   # optical_flowSuper.m:3,19-26 as shipped
        run(n, "super", 11, 3, 16.0, 0.2, 0.75, "driver L=3 K=11")
