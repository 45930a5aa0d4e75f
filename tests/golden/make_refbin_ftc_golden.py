"""Golden vectors for flowToColor_mex produced by the REFERENCE'S OWN BINARY (flowToColor_mex.mexw64, run through
oracle/refbin/peload_ftc.c).  Run in the build container (needs /root/reference):
    python tests/golden/make_refbin_ftc_golden.py  ->  tests/golden/flow_to_color_refbin.npz
Inputs: 64 x 80 crops of the ten Middlebury ground-truth flows the reference ships (each placed where the file has unknown-flow
pixels, if it has any) and synthetic fields that reach every colour-wheel segment, the wrap-around bin, the saturation branch,
NaNs, unknown markers, signed zeros and the all-zero field (maxrad = 0 -> division by eps)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refbin import refbin

MB = "/root/reference/middlebury"
SEQS = ["Cones", "Dimetrodon", "Grove2", "Grove3", "Hydrangea", "Teddy", "Urban2", "Urban3", "Venus", "rubberwhale"]


def read_flo(path):
    """Middlebury .flo: float32 tag 202021.25, int32 width, height, then rows of interleaved (u, v) float32 (readFlowFile.m)."""
    with open(path, "rb") as f:
        tag = np.fromfile(f, np.float32, 1)[0]
        w, h = np.fromfile(f, np.int32, 2)
        assert tag == np.float32(202021.25)
        d = np.fromfile(f, np.float32, 2 * w * h).reshape(h, w, 2)
    return np.asfortranarray(d.astype(np.float64))


def crop_with_unknowns(flow, ch=64, cw=80):
    unk = (np.abs(flow) > 1e9).any(axis=2)
    M, N = unk.shape
    if unk.any():
        r, c = np.argwhere(unk)[len(np.argwhere(unk)) // 2]
    else:
        r, c = M // 2, N // 2
    r0, c0 = int(np.clip(r - ch // 2, 0, M - ch)), int(np.clip(c - cw // 2, 0, N - cw))
    return np.asfortranarray(flow[r0:r0 + ch, c0:c0 + cw, :])


def cases():
    for s in SEQS:
        yield "mb_" + s, crop_with_unknowns(read_flo(os.path.join(MB, s, "flow10.flo")))
    rng = np.random.default_rng(20181003)
    M, N = 40, 56
    for scale in (0.01, 1.0, 30.0):
        yield "normal_%g" % scale, np.asfortranarray(rng.normal(0, scale, (M, N, 2)))
    yy, xx = np.mgrid[0:M, 0:N]
    ang = 2 * np.pi * (xx + N * yy) / (M * N)                    # every direction, radius 0.015..1 of the maximum after normalisation
    rad = 5.0 * (0.02 + 1.28 * yy / (M - 1))
    f = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=2)
    f[0, 0] = (6.5, 0.0)
    yield "wheel", np.asfortranarray(f)
    yield "zeros", np.zeros((7, 9, 2), order="F")
    f = np.asfortranarray(rng.normal(0, 2.0, (M, N, 2)))
    f[3, 4, 0] = np.nan
    f[5, 6, 1] = np.nan
    f[7, 8, :] = np.nan
    f[9, 1, 0] = 1.666666752e9                                   # the marker value in the Middlebury files
    f[10, 2, 1] = -2e9
    f[11, 3, :] = (1e9, -1e9)                                    # exactly the threshold: NOT unknown (strict >)
    yield "nan_unknown", f
    f = np.zeros((6, 8, 2), order="F")
    f[..., 0] = np.array([0.0, -0.0, 1.0, -1.0, 0.0, -0.0, 3.0, -3.0])[None, :]
    f[..., 1] = np.array([0.0, -0.0, 0.0, -0.0, 2.0, -2.0][:6])[:, None]
    yield "axes_signed_zero", f
    yield "constant", np.asfortranarray(np.broadcast_to(np.array([1.25, -0.75]), (5, 6, 2)).copy())
    yield "single_pixel", np.asfortranarray(np.array([[[0.3, -4.0]]]))
    yield "column_vector", np.asfortranarray(rng.normal(0, 1.0, (37, 1, 2)))
    yield "two_rows", np.asfortranarray(rng.normal(0, 1.0, (2, 37, 2)))
    # (a 1 x N field with N > 1 is REJECTED by the binary: MATLAB Coder's run-time check on the vector index `tmp(k0)` of
    #  computeColor.m:57 raises Coder:FE:PotentialMatrixMatrix; the .m file fails there too -- the column it gets back no longer
    #  conforms with the 1 x N arrays around it; the oracle and the product treat 1 x N like any other field)


if __name__ == "__main__":
    out = {}
    n = 0
    for name, flow in cases():
        img, flo, minu, maxu, minv, maxv, unk = refbin.flowToColor_mex(flow)
        out.update({name + "__flow": flow, name + "__img": img, name + "__flo": flo, name + "__range": np.array([minu, maxu, minv, maxv]),
                    name + "__unknown": unk})
        n += 1
    path = os.path.join(HERE, "flow_to_color_refbin.npz")
    np.savez_compressed(path, **out)
    print("wrote", n, "cases,", os.path.getsize(path) // 1024, "KiB")
