"""Coarse-to-fine probe (development aid): synthetic large-displacement pair and, when data/_middlebury exists, Urban2/Urban3."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("gqmap-opticalflow_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M, N = 160, 224
I1, I2, flow, r = pkg.synthetic_pair(M, N, seed=5, flow_scale=3.0)
opts = dict(K=5, its=int(os.environ.get("CTF_ITS", 1500)), epsn=1e-6, lambdas=5.0, lambdad=1.0)
t0 = time.time()
warp, levels = pkg.optical_flow_ctf(I1, I2, flow, opts, scales=(1 / 4, 1 / 2, 1), seed=1, verbose=True)
mu, sigma, rou, AEPE, Energy = pkg.gqmap_ctf(opts, I1, I2, flow, seed=1)
print("synthetic x3: ctf final", levels[-1]["aepe_after"], "single-scale", np.nanmax(AEPE), "%.1fs" % (time.time() - t0), flush=True)
for name in sys.argv[1:]:
    d = np.load(os.path.join(ROOT, "data", "_middlebury", name + ".npz"))
    g1, g2 = pkg.rgb2gray(d["frame10"]).astype(np.float64), pkg.rgb2gray(d["frame11"]).astype(np.float64)
    img, tflow, minu, maxu, minv, maxv, unk = pkg.flowToColor_mex(np.asfortranarray(d["flow10"].astype(np.float64)))
    o = dict(K=11, its=int(os.environ.get("CTF_ITS", 3000)), epsn=0.001 ** 2, lambdas=5.0, lambdad=1.0)   # legacy/optical_flow_ctf.m:13-17
    t0 = time.time()
    warp, levels = pkg.optical_flow_ctf(g1, g2, tflow, o, seed=1)
    err = np.sqrt(((tflow - warp) ** 2).sum(axis=2))
    err[unk] = 0
    print(name, "ctf AEPE per level:", ["%.3f->%.3f" % (l["aepe_before"], l["aepe_after"]) for l in levels],
          "final AEPE (unknown zeroed, interior) %.3f" % err[1:-1, 1:-1].mean(), "%.1fs" % (time.time() - t0), flush=True)
