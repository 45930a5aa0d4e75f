"""BASELINE configs[0] on the real frame pair: Middlebury RubberWhale (full 388x584 frame from data/_middlebury, else the 96x128 crop
of tests/golden), L=1 Gaussian, K=3 (3x3 Gauss-Hermite): fp64 CPU oracle vs the same oracle started one fp32 rounding away vs the CUDA
path, all from the SAME initial state.  Reports per-iteration Energy agreement while the trajectories are still together, AEPE against
the ground truth along the run and the end-point distance between the final flows.  usage: epe_study_real.py [its] [crop]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as O
pkg = importlib.import_module("gqmap-opticalflow_b200")
its = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
full = os.path.join(ROOT, "data", "_middlebury", "RubberWhale.npz")
if os.path.exists(full) and (len(sys.argv) < 3 or sys.argv[2] != "crop"):
    d = np.load(full)
    I1, I2 = pkg.rgb2gray(d["frame10"]).astype(np.float64), pkg.rgb2gray(d["frame11"]).astype(np.float64)
    raw = np.asfortranarray(d["flow10"].astype(np.float64))
    what = "RubberWhale full frame"
else:
    d = np.load(os.path.join(ROOT, "tests", "golden", "rubberwhale_crop.npz"))
    I1, I2, raw = d["I1"].astype(np.float64), d["I2"].astype(np.float64), np.asfortranarray(d["flow"].astype(np.float64))
    what = "RubberWhale 96x128 crop"
I1, I2 = np.asfortranarray(I1), np.asfortranarray(I2)
img, tflow, minu, maxu, minv, maxv, unk = pkg.flowToColor_mex(raw)
M, N = I1.shape
L, K = 1, 3
cfg = O.make_config(M, N, L, K, lambdas=5.0, epsn=1e-6, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
st = O.init_state(cfg, 11)
for f in ("muu", "muv", "sigu", "sigv"):
    getattr(st, f)[...] = getattr(st, f).astype(np.float32).astype(np.float64)
VV = O.get_vv(I2)
marks = sorted(set([1, 10, 30, 100, 300, 1000] + list(range(3000, its + 1, 3000)) + [its]))
marks = [m for m in marks if m <= its]
def aepe_of(muu, muv):
    flow = np.dstack([muu[:, :, 0], muv[:, :, 0]]).copy()
    flow[np.repeat(unk[:, :, None], 2, axis=2)] = 0
    return float(np.sqrt(((tflow - flow) ** 2).sum(axis=2))[1:-1, 1:-1].mean())
def run_oracle(s0):
    s = s0.copy(); out = {}; it = 1; E_all = []
    for m in marks:
        _, _, _, E, dm, ds = O.run(cfg, I1, VV, s, it, 10 ** 6, m - it + 1)
        E_all.append(E); it = m + 1
        out[m] = (aepe_of(s.muu, s.muv), s.muu.copy(), s.muv.copy())
    return out, np.concatenate(E_all)
t0 = time.time(); ref, E_ref = run_oracle(st); t_cpu = time.time() - t0
per0 = st.copy(); per0.muu *= (1 + 2.0 ** -24)
per, E_per = run_oracle(per0)
opts = dict(K=K, L=L, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
gpu = {}; E_gpu = []; t_gpu = 0.0
with pkg.Solver(opts, I1, I2) as s:
    s.set_state(dict(muu=st.muu, muv=st.muv, sigmau=st.sigu, sigmav=st.sigv, pn=st.pn, rou=st.rou, w=st.w), T=0.0)
    it = 1
    for m in marks:
        r = s.step(m - it + 1); it = m + 1; t_gpu += r["ms"] / 1e3; E_gpu.append(r["Energy"])
        g = s.get_state(); gpu[m] = (aepe_of(g["muu"], g["muv"]), g["muu"], g["muv"])
E_gpu = np.concatenate(E_gpu)
dist = lambda a, b: float(np.sqrt((a[1] - b[1]) ** 2 + (a[2] - b[2]) ** 2)[1:-1, 1:-1].mean())
print("%s %dx%d, L=%d K=%d (BASELINE configs[0]), %d iterations; oracle %.1f s on %d host threads, CUDA %.3f s" % (what, M, N, L, K, its, t_cpu, os.cpu_count(), t_gpu))
print("  per-iteration Energy, relative difference to the oracle:  it      CUDA      perturbed oracle")
for i in (1, 2, 3, 5, 10, 20, 30, 50, 100, 300, 1000, its):
    if i <= its:
        print("      %28s %6d  %9.2e  %9.2e" % ("", i, abs(E_gpu[i - 1] / E_ref[i - 1] - 1), abs(E_per[i - 1] / E_ref[i - 1] - 1)))
print("  AEPE vs ground truth (and mean end-point distance to the oracle's flow) at iteration:")
for m in marks:
    print("      it %6d  oracle %.4f | perturbed oracle %.4f (dist %.2e) | CUDA %.4f (dist %.2e)" % (
        m, ref[m][0], per[m][0], dist(per[m], ref[m]), gpu[m][0], dist(gpu[m], ref[m])))
