"""Generates the committed golden fixtures under tests/golden/ from the CPU fp64 oracle (and, for the Middlebury crop, from
the reference's own data files under /root/reference, which do not travel to the GPU box).  Re-run only when the oracle
changes on purpose:   python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O                                     # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "gqmap-opticalflow_b200"))
import frames                                                      # noqa: E402  (no libqgmap needed)

REF = "/root/reference"


def f32(a):
    return a.astype(np.float32).astype(np.float64)


def problem(Mo, No, L, K, sup, seed, T, warm):
    """Reference-style problem advanced `warm` oracle iterations so the state is a realistic mid-trajectory one."""
    I1, I2, flow, (minu, maxu, minv, maxv) = frames.synthetic_pair(Mo, No, seed=seed)
    cfg = O.make_config(Mo, No, L, K, super=sup, lambdas=16.0 if sup else 5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv,
                        drate=0.75)
    st = O.init_state(cfg, seed + 1, T=T)
    VV = O.get_vv(I2)
    if warm:
        O.run(cfg, I1, VV, st, 1, 10 ** 6, warm)
    for f in ("muu", "muv", "sigu", "sigv", "pn", "rou"):
        getattr(st, f)[...] = f32(getattr(st, f))                 # fp32-representable: the device keeps beliefs in fp32
    return cfg, I1, I2, VV, flow, st


def steps_fixture(name, Mo, No, L, K, sup, seed, T, warm, nsteps=2):
    cfg, I1, I2, VV, flow, st = problem(Mo, No, L, K, sup, seed, T, warm)
    it0 = warm + 1
    g = O.gradients(cfg, I1, VV, st, assemble=True)
    out = dict(Mo=Mo, No=No, L=L, K=K, super=int(sup), T=st.T, it0=it0, I1=I1, I2=I2, flow=flow,
               minu=cfg.minu, maxu=cfg.maxu, minv=cfg.minv, maxv=cfg.maxv, lambdas=cfg.lambdas, drate=cfg.drate,
               muu=st.muu, muv=st.muv, sigu=st.sigu, sigv=st.sigv, pn=st.pn, rou=st.rou, w=st.w, alpha=st.alpha,
               G_muu=g["dmuu"], G_muv=g["dmuv"], G_sigu=g["dsigmau"], G_sigv=g["dsigmav"], dpn=g["dpn"], drou=g["drou"],
               e_px=g["nEnergy"] + g["eEnergy"].sum(axis=(3, 4)), da_px=g["dan"] + g["dae"].sum(axis=(3, 4)), dalpha=g["dalpha"])
    # single steps, each from the (fp32-rounded) state the previous one produced
    E, dm, ds = [], [], []
    cur = st.copy()
    states = []
    for k in range(nsteps):
        _, _, _, e, a, b = O.run(cfg, I1, VV, cur, it0 + k, 10 ** 6, 1)
        E.append(e[0]); dm.append(a[0]); ds.append(b[0])
        for f in ("muu", "muv", "sigu", "sigv", "pn", "rou"):
            getattr(cur, f)[...] = f32(getattr(cur, f))
        states.append({f: getattr(cur, f).copy() for f in ("muu", "muv", "sigu", "sigv", "pn", "rou")})
    out.update(step_Energy=np.array(E), step_ptdmu=np.array(dm), step_ptdsigma=np.array(ds))
    for k, s in enumerate(states):
        for f, a in s.items():
            out["s%d_%s" % (k, f)] = a.astype(np.float32)
    m = O.find_map(st.alpha, st.muu, st.sigu, st.muv, st.sigv)
    out.update(map=m, logp=O.profile_logp(cfg, I1, VV, m))
    unk = np.zeros((Mo, No), bool)
    unk[3:9, 5:11] = True
    out.update(unknown=unk, aepe=O.aepe(cfg, m, flow, unk))
    # beliefs are fp32-representable by construction and gradients are compared at fp32 accuracy: store them as float32
    for k in ("muu", "muv", "sigu", "sigv", "pn", "rou", "G_muu", "G_muv", "G_sigu", "G_sigv", "dpn", "drou", "e_px", "da_px"):
        out[k] = np.asarray(out[k], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: np.asarray(v).shape for k, v in out.items() if np.asarray(v).ndim > 1 and k in ("I1", "muu", "rou", "map")},
          "E", out["step_Energy"])


def middlebury_fixture():
    """RubberWhale crop (BASELINE configs[0]: L=1, K=3 on the real frames): grey frames via MATLAB-equivalent rgb2gray,
    ground-truth flow via readFlowFile + flowToColor (unknown mask, clamp range), oracle energies of the first steps."""
    from PIL import Image
    d = os.path.join(REF, "middlebury", "rubberwhale")
    g1 = frames.rgb2gray(np.asarray(Image.open(os.path.join(d, "frame10.png"))))
    g2 = frames.rgb2gray(np.asarray(Image.open(os.path.join(d, "frame11.png"))))
    gt = frames.readFlowFile(os.path.join(d, "flow10.flo"))
    r0, c0, Mo, No = 150, 250, 96, 128
    I1 = np.asfortranarray(g1[r0:r0 + Mo, c0:c0 + No].astype(np.float64))
    I2 = np.asfortranarray(g2[r0:r0 + Mo, c0:c0 + No].astype(np.float64))
    flow = np.asfortranarray(gt[r0:r0 + Mo, c0:c0 + No].copy())
    flow[10:14, 20:26] = 1.666666752e9                                  # Middlebury's "unknown" marker
    img, flo, minu, maxu, minv, maxv, unk = O.flow_to_color(flow)
    full_stats = O.flow_to_color(gt)[2:6]
    cfg = O.make_config(Mo, No, 1, 3, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
    st = O.init_state(cfg, 2024)
    for f in ("muu", "muv", "sigu", "sigv"):
        getattr(st, f)[...] = f32(getattr(st, f))
    VV = O.get_vv(I2)
    ref = st.copy()
    _, _, _, E, dm, ds = O.run(cfg, I1, VV, ref, 1, 10 ** 6, 2)
    np.savez_compressed(os.path.join(HERE, "rubberwhale_crop.npz"), I1=I1.astype(np.uint8), I2=I2.astype(np.uint8), flow=flow.astype(np.float32),
                        color=img, flo=flo.astype(np.float32), stats=np.array([minu, maxu, minv, maxv]), unknown=unk,
                        full_stats=np.array(full_stats), full_shape=np.array(g1.shape),
                        grey_checksum=np.array([int(g1.astype(np.int64).sum()), int(g2.astype(np.int64).sum())]),
                        muu=st.muu.astype(np.float32), muv=st.muv.astype(np.float32), sigu=st.sigu.astype(np.float32),
                        sigv=st.sigv.astype(np.float32), w=st.w, Energy=E, ptdmu=dm, ptdsigma=ds)
    print("rubberwhale_crop", I1.shape, "stats", minu, maxu, minv, maxv, "full", full_stats, "E", E)


def tables_fixture():
    out = {}
    for n in (2, 3, 5, 9, 11, 17):
        x, w = O.gauss_hermite(n)
        out["x%d" % n], out["w%d" % n] = x, w
    np.savez_compressed(os.path.join(HERE, "gauss_hermite.npz"), **out)


if __name__ == "__main__":
    steps_fixture("full_L2K5.npz", 40, 56, 2, 5, False, 11, 0.0, warm=25)
    steps_fixture("full_L3K4_T.npz", 36, 44, 3, 4, False, 21, 0.3, warm=8)
    steps_fixture("super_L3K3.npz", 64, 96, 3, 3, True, 31, 0.2, warm=25)
    tables_fixture()
    if os.path.isdir(REF):
        middlebury_fixture()
