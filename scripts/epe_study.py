"""End-point-error study (north_star gate: final flow EPE within 1e-3 px of the reference's CPU path).
Runs the fp64 oracle, an fp32-perturbed copy of the oracle, and the CUDA path from the SAME initial state for `its`
iterations on a small synthetic pair and reports AEPE (vs ground truth) and the mean endpoint distance between the MAP flows.
usage: epe_study.py [M N L K its variant]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
pkg = importlib.import_module("gqmap-opticalflow_b200")
a = sys.argv[1:]
M = int(a[0]) if len(a) > 0 else 64
N = int(a[1]) if len(a) > 1 else 96
L = int(a[2]) if len(a) > 2 else 2
K = int(a[3]) if len(a) > 3 else 3
its = int(a[4]) if len(a) > 4 else 3000
variant = a[5] if len(a) > 5 else "full"
sup = variant == "super"
I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N, seed=7)
# integer-valued frames like double(rgb2gray(uint8)) in the reference driver (optical_flow.m:10-11)
I1, I2 = np.asfortranarray(np.round(I1)), np.asfortranarray(np.round(I2))
cfg = O.make_config(M, N, L, K, super=sup, lambdas=16.0 if sup else 5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, drate=0.75)
T0 = 0.2 if sup else 0.0
st = O.init_state(cfg, 11, T=T0)
for f in ("muu", "muv", "sigu", "sigv"):
    getattr(st, f)[...] = getattr(st, f).astype(np.float32).astype(np.float64)
VV = O.get_vv(I2)
unk = np.zeros((M, N), bool)
def finish(s_alpha, muu, sigu, muv, sigv):
    m = O.find_map(s_alpha, muu, sigu, muv, sigv)
    return m, O.aepe(cfg, m, flow, unk)
t0 = time.time()
ref = st.copy(); O.run(cfg, I1, VV, ref, 1, its, its)
m_ref, a_ref = finish(ref.alpha, ref.muu, ref.sigu, ref.muv, ref.sigv)
per = st.copy(); per.muu *= (1 + 2.0 ** -24); O.run(cfg, I1, VV, per, 1, its, its)
m_per, a_per = finish(per.alpha, per.muu, per.sigu, per.muv, per.sigv)
t_cpu = time.time() - t0
opts = dict(K=K, L=L, temperature=T0, drate=0.75, epsn=cfg.epsn, lambdad=1.0, lambdas=cfg.lambdas, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
with pkg.Solver(opts, I1, I2, variant=variant) as s:
    s.set_state(dict(muu=st.muu, muv=st.muv, sigmau=st.sigu, sigmav=st.sigv, pn=st.pn, rou=st.rou, w=st.w), T=T0)
    r = s.step(its, its=its)
    g = s.get_state()
m_gpu, a_gpu = finish(g["alpha"], g["muu"], g["sigmau"], g["muv"], g["sigmav"])
sc = 4 if sup else 1
dist = lambda x, y: float(np.sqrt(((x - y) ** 2).sum(axis=2))[1:-1, 1:-1].mean())
print("%s %dx%d L=%d K=%d its=%d (cpu %.0fs)" % (variant, M, N, L, K, its, t_cpu))
print("  AEPE vs ground truth : oracle %.5f | oracle perturbed by 2^-24 %.5f | CUDA %.5f" % (a_ref, a_per, a_gpu))
print("  |AEPE - AEPE_oracle|  : perturbed oracle %.2e | CUDA %.2e" % (abs(a_per - a_ref), abs(a_gpu - a_ref)))
print("  mean endpoint distance between MAP flows: perturbed oracle vs oracle %.2e | CUDA vs oracle %.2e" % (dist(m_per, m_ref), dist(m_gpu, m_ref)))
print("  median / 90%% / max per-pixel distance CUDA vs oracle: %s" % np.percentile(np.sqrt(((m_gpu - m_ref) ** 2).sum(axis=2)), [50, 90, 100]))
print("  median / 90%% / max per-pixel distance perturbed vs oracle: %s" % np.percentile(np.sqrt(((m_per - m_ref) ** 2).sum(axis=2)), [50, 90, 100]))
