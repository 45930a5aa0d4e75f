"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE: gqmap_gpu_mixture.m and gqmap_gpuSuper_mix_entropy.m are run,
unmodified and read from /root/reference, by the mini-MATLAB interpreter oracle/mlab/minimat.py; their MEX calls (get_map_mex,
flowToColor_mex) go to the reference's own .mexw64 machine code through oracle/refbin.  Run in the build container:
    python tests/golden/make_refsrc_golden.py        ->  tests/golden/refsrc_<case>.npz   (a few minutes)
Every file holds the inputs, the `rand` draws the program consumed (so the oracle can start from the same state), the six outputs
of the solver, the state it does not return (pn, rou, w -- read from the function workspace when it ends) and per-iteration probes
taken at the solver's own fprintf (state after the update, alpha, T, Energy(it), mean|dmu|, mean|dsigma|)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference"

# name -> (solver file, Mo, No, L, K, T, drate, lambdas, its, iterations to probe, seed)
CASES = {
    "full_L2K3":      ("gqmap_gpu_mixture", 8, 9, 2, 3, 0.0, 0.5, 5.0, 4, (1, 2, 3, 4), 11),
    "full_L1K4":      ("gqmap_gpu_mixture", 7, 8, 1, 4, 0.0, 0.5, 5.0, 3, (1, 2, 3), 12),
    "full_L3K3_T":    ("gqmap_gpu_mixture", 8, 8, 3, 3, 0.3, 0.5, 5.0, 3, (1, 2, 3), 13),
    "full_alpha":     ("gqmap_gpu_mixture", 5, 5, 2, 3, 0.0, 0.5, 5.0, 504, (1, 499, 500, 501, 502, 503, 504), 14),
    "super_L2K3_T":   ("gqmap_gpuSuper_mix_entropy", 16, 20, 2, 3, 0.2, 0.75, 16.0, 3, (1, 2, 3), 15),
    "super_anneal":   ("gqmap_gpuSuper_mix_entropy", 12, 12, 2, 2, 0.2, 0.75, 16.0, 502, (1, 498, 499, 500, 501, 502), 16),
    "super_L1K3":     ("gqmap_gpuSuper_mix_entropy", 16, 16, 1, 3, 0.2, 0.75, 16.0, 2, (1, 2), 17),
    "full_zero_v":    ("gqmap_gpu_mixture", 8, 8, 2, 3, 0.0, 0.5, 5.0, 3, (1, 2, 3), 18),      # Venus/Teddy/Cones: minv = maxv = 0
    "full_L2K5":      ("gqmap_gpu_mixture", 7, 7, 2, 5, 0.1, 0.5, 5.0, 2, (1, 2), 19),
    # K >= 7 runs the clamp-first sample path of the CUDA kernel; on images this small most quadrature points are clamped to the border
    "full_L2K9":      ("gqmap_gpu_mixture", 6, 7, 2, 9, 0.0, 0.5, 5.0, 2, (1, 2), 20),
    "full_L1K7":      ("gqmap_gpu_mixture", 9, 6, 1, 7, 0.2, 0.5, 5.0, 2, (1, 2), 21),
}
RANGE = dict(minu=-3.0, maxu=2.0, minv=-1.5, maxv=4.0)
CASE_RANGE = {"full_zero_v": dict(minu=-9.375, maxu=7.0, minv=0.0, maxv=0.0)}                    # the Venus ground-truth range
PROBE_FIELDS = ("muu", "muv", "sigmau", "sigmav", "pn", "rou", "w", "alpha")


def frames(Mo, No, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:Mo, 0:No]
    I1 = 127 + 60 * np.sin(xx / 2.1) * np.cos(yy / 2.3) + 40 * rng.random((Mo, No))
    I2 = 127 + 60 * np.sin((xx - 0.8) / 2.1) * np.cos((yy + 0.5) / 2.3) + 40 * rng.random((Mo, No))
    tflow = rng.normal(0, 1.0, (Mo, No, 2))
    unk = rng.random((Mo, No)) < 0.05
    return np.asfortranarray(I1), np.asfortranarray(I2), np.asfortranarray(tflow), np.asfortranarray(unk)


def run_case(name, its=None, probes=None):
    """-> dict of arrays (what the .npz holds)"""
    from oracle.mlab.minimat import Interp
    from oracle.refbin import refbin
    solver, Mo, No, L, K, T, drate, lambdas, its0, probes0, seed = CASES[name]
    its = its0 if its is None else its
    probes = probes0 if probes is None else probes
    I1, I2, tflow, unk = frames(Mo, No, seed)
    rg = CASE_RANGE.get(name, RANGE)
    opts = dict(trueFlow=tflow, unknownIdx=unk, its=float(its), K=float(K), L=float(L), temperature=T, drate=drate, epsn=1e-6,
                lambdad=1.0, lambdas=lambdas, dir="/nonexistent", **rg)
    rng = np.random.default_rng(seed + 1000)
    draws, snaps = [], {}

    def rand(shape):
        a = rng.random(int(np.prod(shape))).reshape(shape, order="F")
        draws.append(np.array(a))
        return a

    def probe(ws):
        it = int(ws["it"])
        if it in probes:
            d = {f: np.array(np.asarray(ws[f], dtype=np.float64), order="F") for f in PROBE_FIELDS}
            d.update(T=float(ws["T"]), Energy=float(np.asarray(ws["Energy"]).reshape(-1, order="F")[it - 1]), ptdmu=float(ws["ptdmu"]),
                     ptdsigma=float(ws["ptdsigma"]), step=float(ws["step"]))
            snaps[it] = d
    interp = Interp([REF], rand=rand, on_fprintf=probe,
                    externals={"get_map_mex": lambda n, *a: (refbin.get_map_mex(*a),),
                               "flowToColor_mex": lambda n, *a: refbin.flowToColor_mex(*a)[:max(n, 1)]})
    mu, sigma, alpha, AEPE, Energy, logP = interp.call(solver, opts, I1, I2, nargout=6)
    ws = interp.last_workspace
    out = dict(I1=I1, I2=I2, tflow=tflow, unknown=unk, mu=mu, sigma=sigma, alpha=np.ravel(alpha), AEPE=np.ravel(AEPE),
               Energy=np.ravel(Energy), logP=np.ravel(logP), pn=np.asarray(ws["pn"]), rou=np.asarray(ws["rou"]), w=np.ravel(ws["w"]),
               it_end=np.array(int(ws["it"])), T_end=np.array(float(ws["T"])),
               meta=np.array([Mo, No, L, K, T, drate, lambdas, its, rg["minu"], rg["maxu"], rg["minv"], rg["maxv"]]),
               solver=np.array(solver), probes=np.array(sorted(snaps)))
    for i, d in enumerate(draws):
        out["draw%d" % i] = d
    for it, d in snaps.items():
        for k, v in d.items():
            out["p%d_%s" % (it, k)] = np.asarray(v)
    return out


RW_SEED, RW_ITS, RW_STRIDE = 2018, 3, 4


def rubberwhale_inputs():
    """BASELINE configs[0] on the reference's own data: the 96 x 128 crop of the RubberWhale pair committed in rubberwhale_crop.npz
    (grey frames, ground truth, unknown mask, clamp range of the whole sequence), L=1, K=3, driver constants of optical_flow.m."""
    z = np.load(os.path.join(HERE, "rubberwhale_crop.npz"))
    I1, I2 = np.asfortranarray(z["I1"].astype(np.float64)), np.asfortranarray(z["I2"].astype(np.float64))
    tflow = np.asfortranarray(z["flo"].astype(np.float64))                   # flowToColor_mex's second output (unknowns zeroed)
    minu, maxu, minv, maxv = (float(x) for x in z["full_stats"])
    opts = dict(trueFlow=tflow, unknownIdx=np.asfortranarray(z["unknown"]), its=float(RW_ITS), K=3.0, L=1.0, temperature=0.0, drate=0.5,
                epsn=0.001 ** 2, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, dir="/nonexistent")
    return I1, I2, opts


def rubberwhale_draws():
    rng = np.random.default_rng(RW_SEED)
    return [rng.random(n).reshape(shp, order="F") for n, shp in ((1, (1, 1)),) + ((96 * 128, (96, 128)),) * 4]


def run_rubberwhale():
    from oracle.mlab.minimat import Interp
    from oracle.refbin import refbin
    I1, I2, opts = rubberwhale_inputs()
    draws = rubberwhale_draws()
    queue, snaps = list(draws), {}

    def rand(shape):
        a = queue.pop(0)
        assert a.size == int(np.prod(shape))
        return a.reshape(shape, order="F")

    def probe(ws):
        it = int(ws["it"])
        d = {f: np.asarray(ws[f], dtype=np.float64) for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou")}
        snaps[it] = dict(Energy=float(np.asarray(ws["Energy"]).reshape(-1, order="F")[it - 1]), ptdmu=float(ws["ptdmu"]),
                         ptdsigma=float(ws["ptdsigma"]), sums=np.array([d[f].sum() for f in d] + [(d[f] ** 2).sum() for f in d]),
                         **{f: np.array(d[f][::RW_STRIDE, ::RW_STRIDE]) for f in d})
    interp = Interp([REF], rand=rand, on_fprintf=probe,
                    externals={"get_map_mex": lambda n, *a: (refbin.get_map_mex(*a),),
                               "flowToColor_mex": lambda n, *a: refbin.flowToColor_mex(*a)[:max(n, 1)]})
    mu, sigma, alpha, AEPE, Energy, logP = interp.call("gqmap_gpu_mixture", opts, I1, I2, nargout=6)
    out = dict(AEPE=np.ravel(AEPE), Energy=np.ravel(Energy), logP=np.ravel(logP), alpha=np.ravel(alpha), seed=np.array(RW_SEED),
               draws_checksum=np.array([d.sum() for d in draws]))
    for it, d in snaps.items():
        for k, v in d.items():
            out["p%d_%s" % (it, k)] = np.asarray(v)
    return out


def run_rubberwhale_full():
    """BASELINE configs[0] at FULL size: the whole 388 x 584 RubberWhale pair (MATLAB rgb2gray of the shipped PNGs), L=1, K=3, one
    iteration of gqmap_gpu_mixture.m incl. its monitoring -- about ten minutes under the interpreter.  The grey frames are stored
    (uint8) so that the GPU box can repeat the step; the ground truth is used for AEPE(1) here but not stored."""
    from PIL import Image
    from oracle.mlab.minimat import Interp
    from oracle.refbin import refbin
    d = os.path.join(REF, "middlebury", "rubberwhale")

    def grey(fn):                                                            # MATLAB rgb2gray on uint8: weighted sum, rounded
        rgb = np.asarray(Image.open(os.path.join(d, fn)).convert("RGB")).astype(np.float64)
        return np.floor(rgb[..., 0] * 0.298936021293775 + rgb[..., 1] * 0.587043074451121 + rgb[..., 2] * 0.114020904255103 + 0.5).astype(np.uint8)
    g1, g2 = grey("frame10.png"), grey("frame11.png")
    z = np.load(os.path.join(HERE, "rubberwhale_crop.npz"))
    assert [int(g1.astype(np.int64).sum()), int(g2.astype(np.int64).sum())] == list(z["grey_checksum"])       # same conversion as the crop fixture
    Mo, No = g1.shape
    with open(os.path.join(d, "flow10.flo"), "rb") as f:
        f.read(12)
        gt = np.fromfile(f, np.float32).reshape(Mo, No, 2).astype(np.float64)
    _, tflow, minu, maxu, minv, maxv, unk = refbin.flowToColor_mex(np.asfortranarray(gt))                      # optical_flow.m:12-13, the binary
    I1, I2 = np.asfortranarray(g1.astype(np.float64)), np.asfortranarray(g2.astype(np.float64))
    opts = dict(trueFlow=tflow, unknownIdx=unk, its=1.0, K=3.0, L=1.0, temperature=0.0, drate=0.5, epsn=0.001 ** 2, lambdad=1.0,
                lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, dir="/nonexistent")
    rng = np.random.default_rng(RW_SEED + 1)
    draws = [rng.random(n).reshape(shp, order="F") for n, shp in ((1, (1, 1)),) + ((Mo * No, (Mo, No)),) * 4]
    queue, snap = list(draws), {}

    def rand(shape):
        return queue.pop(0).reshape(shape, order="F")

    def probe(ws):
        dd = {f: np.asarray(ws[f], dtype=np.float64) for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou")}
        snap.update(ptdmu=float(ws["ptdmu"]), ptdsigma=float(ws["ptdsigma"]),
                    sums=np.array([dd[f].sum() for f in dd] + [(dd[f] ** 2).sum() for f in dd]),
                    **{f: np.array(dd[f][::8, ::8]) for f in dd})
    interp = Interp([REF], rand=rand, on_fprintf=probe,
                    externals={"get_map_mex": lambda n, *a: (refbin.get_map_mex(*a),),
                               "flowToColor_mex": lambda n, *a: refbin.flowToColor_mex(*a)[:max(n, 1)]})
    mu, sigma, alpha, AEPE, Energy, logP = interp.call("gqmap_gpu_mixture", opts, I1, I2, nargout=6)
    out = dict(I1=g1, I2=g2, range=np.array([minu, maxu, minv, maxv]), AEPE=np.ravel(AEPE), Energy=np.ravel(Energy), logP=np.ravel(logP),
               seed=np.array(RW_SEED + 1), draws_checksum=np.array([x.sum() for x in draws]))
    out.update({"p1_" + k: np.asarray(v) for k, v in snap.items()})
    return out


def _middlebury_step(seq, solver, L, K, consts, seed, stride, window=None, keep_truth=False, its=1, full_at=()):
    """`its` iterations of `solver` (.m source, executed) on the MATLAB-rgb2gray frames of a shipped sequence, or on a window
    (r0, r1, c0, c1) of it with the clamp range of the WHOLE sequence (what the driver computes, optical_flow.m:12-13).  Stores the
    grey frames (uint8) so that a box without the reference tree can repeat the step, probes of the state at the solver's own
    fprintf (strided; the WHOLE state, incl. alpha and w, at the iterations in full_at: keys f<it>_<field>), and the ground truth of a
    window (float32, as in the .flo file) when asked."""
    from PIL import Image
    from oracle.mlab.minimat import Interp
    from oracle.refbin import refbin
    d = os.path.join(REF, "middlebury", seq)

    def grey(fn):                                                            # MATLAB rgb2gray on uint8: weighted sum, rounded
        rgb = np.asarray(Image.open(os.path.join(d, fn)).convert("RGB")).astype(np.float64)
        return np.floor(rgb[..., 0] * 0.298936021293775 + rgb[..., 1] * 0.587043074451121 + rgb[..., 2] * 0.114020904255103 + 0.5).astype(np.uint8)
    g1, g2 = grey("frame10.png"), grey("frame11.png")
    Mo, No = g1.shape
    with open(os.path.join(d, "flow10.flo"), "rb") as f:
        f.read(12)
        gt = np.fromfile(f, np.float32).reshape(Mo, No, 2).astype(np.float64)
    _, tflow, minu, maxu, minv, maxv, unk = refbin.flowToColor_mex(np.asfortranarray(gt))                      # optical_flow.m:12-13, the binary
    if window:
        r0, r1, c0, c1 = window
        g1, g2 = g1[r0:r1, c0:c1], g2[r0:r1, c0:c1]
        tflow, unk = np.asfortranarray(tflow[r0:r1, c0:c1]), np.asfortranarray(np.asarray(unk)[r0:r1, c0:c1])
        Mo, No = g1.shape
    sup = solver == "gqmap_gpuSuper_mix_entropy"
    M, N = (Mo // 4, No // 4) if sup else (Mo, No)
    I1, I2 = np.asfortranarray(g1.astype(np.float64)), np.asfortranarray(g2.astype(np.float64))
    opts = dict(trueFlow=tflow, unknownIdx=unk, its=float(its), K=float(K), L=float(L), epsn=0.001 ** 2, lambdad=1.0,
                minu=minu, maxu=maxu, minv=minv, maxv=maxv, dir="/nonexistent", **consts)
    rng = np.random.default_rng(seed)
    draws = [rng.random(n).reshape(shp, order="F") for n, shp in ((L, (1, 1, L)),) + ((M * N * L, (M, N, L)),) * 4]   # :18-22
    queue, snaps, fulls = list(draws), {}, {}

    def rand(shape):
        a = queue.pop(0)
        assert a.size == int(np.prod(shape))
        return a.reshape(shape, order="F")

    def probe(ws):
        dd = {f: np.asarray(ws[f], dtype=np.float64) for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou")}
        it = int(ws["it"])
        snaps[it] = dict(ptdmu=float(ws["ptdmu"]), ptdsigma=float(ws["ptdsigma"]), T=float(ws["T"]),
                         sums=np.array([dd[f].sum() for f in dd] + [(dd[f] ** 2).sum() for f in dd]),
                         **{f: np.array(dd[f][::stride, ::stride]) for f in dd})
        if it in full_at:
            fulls[it] = dict({f: np.array(dd[f], order="F") for f in dd}, alpha=np.ravel(np.asarray(ws["alpha"], dtype=np.float64)).copy(),
                             w=np.ravel(np.asarray(ws["w"], dtype=np.float64)).copy())
    interp = Interp([REF], rand=rand, on_fprintf=probe,
                    externals={"get_map_mex": lambda n, *a: (refbin.get_map_mex(*a),),
                               "flowToColor_mex": lambda n, *a: refbin.flowToColor_mex(*a)[:max(n, 1)]})
    mu, sigma, alpha, AEPE, Energy, logP = interp.call(solver, opts, I1, I2, nargout=6)
    assert not queue
    out = dict(I1=g1, I2=g2, range=np.array([minu, maxu, minv, maxv]), AEPE=np.ravel(AEPE), Energy=np.ravel(Energy), logP=np.ravel(logP),
               alpha=np.ravel(alpha), seed=np.array(seed), draws_checksum=np.array([x.sum() for x in draws]))
    if keep_truth:
        out.update(tflow=np.asarray(tflow, dtype=np.float32), unknown=np.asarray(unk, dtype=bool))
    for it, snap in snaps.items():
        out.update({"p%d_%s" % (it, k): np.asarray(v) for k, v in snap.items()})
    for it, full in fulls.items():
        out.update({"f%d_%s" % (it, k): np.asarray(v) for k, v in full.items()})
    return out


# Executed-source cases on the reference's OWN data (rgb2gray of the shipped PNGs, ground truth of the shipped .flo files):
# name -> (sequence, solver file, L, K, driver constants, seed, probe stride, window (r0, r1, c0, c1) or None, keep ground truth, its)
FULL_C = dict(temperature=0.0, drate=0.5, lambdas=5.0)                      # optical_flow.m:16-23
SUPER_C = dict(temperature=0.2, drate=0.75, lambdas=16.0)                   # optical_flowSuper.m:16-23
REAL_CASES = {
    # BASELINE configs[2] at FULL size: the whole 480 x 640 Urban2 pair, 120 x 160 x 3 beliefs, 400 node_pot calls each (~80 minutes)
    "urban2_super_full_L3K5": ("Urban2", "gqmap_gpuSuper_mix_entropy", 3, 5, SUPER_C, 2020, 4, None, False, 1),
    # the metric's own instantiation (BASELINE configs[3]/[4]: full resolution, L=3, K=5), two iterations (~25 minutes)
    "grove2_window_L3K5": ("Grove2", "gqmap_gpu_mixture", 3, 5, FULL_C, 2021, 4, (176, 304, 240, 400), True, 2),
    # BASELINE configs[1] (the eight ground-truth sequences, L=2, driver default K=9): a window of Dimetrodon (~20 minutes)
    "dimetrodon_window_L2K9": ("Dimetrodon", "gqmap_gpu_mixture", 2, 9, FULL_C, 2024, 4, (150, 246, 200, 328), True, 1),
    # optical_flow.m AS SHIPPED: Teddy, K=9, L=3; u range [-52.75, 0], v range [0, 0] (sigma_u starts at 53 px: most samples clamp)
    "teddy_window_L3K9": ("Teddy", "gqmap_gpu_mixture", 3, 9, FULL_C, 2022, 4, (120, 216, 160, 288), True, 1),
    # optical_flowSuper.m AS SHIPPED: Venus, K=11, L=3, super-pixel variant; v range [0, 0]
    "venus_super_window_L3K11": ("Venus", "gqmap_gpuSuper_mix_entropy", 3, 11, SUPER_C, 2023, 2, (100, 228, 120, 280), True, 1),
}


# A longer run, for single steps from LATER states of the reference's own trajectory on its own data (at iteration 1 every correlation is
# still zero): a 48 x 64 window of Grove2, L=3, K=5, 16 iterations, the whole state kept after iterations 15 and 16 (~30 minutes)
REAL_LONG = {
    "grove2_window_L3K5_it16": ("Grove2", "gqmap_gpu_mixture", 3, 5, FULL_C, 2025, 4, (208, 256, 280, 344), True, 16, (15, 16)),
    # the super-pixel variant (64 x 80 window of Venus = 16 x 20 beliefs, v range [0, 0]) and the driver's K=9 (32 x 48 window of Dimetrodon, L=2)
    "venus_super_window_L3K5_it12": ("Venus", "gqmap_gpuSuper_mix_entropy", 3, 5, SUPER_C, 2026, 2, (120, 184, 140, 220), True, 12, (11, 12)),
    "dimetrodon_window_L2K9_it8": ("Dimetrodon", "gqmap_gpu_mixture", 2, 9, FULL_C, 2027, 4, (170, 202, 230, 278), True, 8, (7, 8)),
}


def run_real(name, crop=None):
    """crop = (rows, cols): dry runs of the full-size case."""
    if name in REAL_LONG:
        seq, solver, L, K, consts, seed, stride, window, keep, its, full_at = REAL_LONG[name]
        return _middlebury_step(seq, solver, L, K, consts, seed, stride, window=window, keep_truth=keep, its=its, full_at=full_at)
    seq, solver, L, K, consts, seed, stride, window, keep, its = REAL_CASES[name]
    if crop:
        window = (0, crop[0], 0, crop[1])
    return _middlebury_step(seq, solver, L, K, consts, seed, stride, window=window, keep_truth=keep, its=its)


def run_host_io():
    """The host-side .m files of the drivers' path, executed: readFlowFile.m, legacy/writeFlowFile.m and legacy/flowToColor.m +
    legacy/computeColor.m with the optional maxFlow argument (the compiled flowToColor_mex takes none)."""
    import tempfile
    from oracle.mlab.minimat import Interp
    interp = Interp([REF, os.path.join(REF, "legacy")])
    rng = np.random.default_rng(99)
    flow = np.asfortranarray(rng.normal(0, 2.0, (9, 7, 2)).astype(np.float32).astype(np.float64))       # .flo stores float32
    flow[2, 3, :] = 1.666666752e9
    flow[5, 1, 0] = -1.666666752e9
    out = dict(flow=flow)
    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, "a.flo")
        interp.call("writeFlowFile", flow, fn, nargout=0)
        out["flo_bytes"] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
        out["flo_read_back"] = interp.call("readFlowFile", fn)
    for tag, mf in (("mf4", 4.0), ("mf05", 0.5), ("mfneg", -1.0)):
        r = interp.call("flowToColor", flow, mf, nargout=7)
        out.update({tag + "_img": np.asarray(r[0]), tag + "_flo": np.asarray(r[1]), tag + "_range": np.array([float(x) for x in r[2:6]]),
                    tag + "_unknown": np.asarray(r[6], dtype=bool), tag + "_maxflow": np.array(mf)})
    return out


if __name__ == "__main__":
    if not sys.argv[1:] or "host_io" in sys.argv[1:]:
        out = run_host_io()
        np.savez_compressed(os.path.join(HERE, "refsrc_host_io.npz"), **out)
        print("host_io         -> %d KiB" % (os.path.getsize(os.path.join(HERE, "refsrc_host_io.npz")) // 1024), flush=True)
    if not sys.argv[1:] or "rubberwhale" in sys.argv[1:]:
        t = time.time()
        out = run_rubberwhale()
        np.savez_compressed(os.path.join(HERE, "refsrc_rubberwhale_L1K3.npz"), **out)
        print("rubberwhale     %6.1f s  Energy=%s AEPE(1)=%.6f -> %d KiB" % (time.time() - t, out["Energy"], out["AEPE"][0],
                                                                            os.path.getsize(os.path.join(HERE, "refsrc_rubberwhale_L1K3.npz")) // 1024), flush=True)
    if "rubberwhale_full" in sys.argv[1:]:                                  # ~10 minutes: only on request
        t = time.time()
        out = run_rubberwhale_full()
        np.savez_compressed(os.path.join(HERE, "refsrc_rubberwhale_full_L1K3.npz"), **out)
        print("rubberwhale_full %6.1f s  Energy(1)=%.9e AEPE(1)=%.6f logP(1)=%.6e -> %d KiB" % (
            time.time() - t, out["Energy"][0], out["AEPE"][0], out["logP"][0],
            os.path.getsize(os.path.join(HERE, "refsrc_rubberwhale_full_L1K3.npz")) // 1024), flush=True)
    for name in [a for a in sys.argv[1:] if a in REAL_CASES or a in REAL_LONG]:               # 20-80 minutes each: only on request
        t = time.time()
        out = run_real(name)
        path = os.path.join(HERE, "refsrc_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("%s %6.1f s  Energy=%s AEPE(1)=%.6f logP(1)=%.6e -> %d KiB" % (name, time.time() - t, out["Energy"], out["AEPE"][0], out["logP"][0],
                                                                              os.path.getsize(path) // 1024), flush=True)
    for name in ([a for a in sys.argv[1:] if a not in ("host_io", "rubberwhale", "rubberwhale_full") and a not in REAL_CASES and a not in REAL_LONG] or ([] if sys.argv[1:] else CASES)):
        t = time.time()
        out = run_case(name)
        path = os.path.join(HERE, "refsrc_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("%-14s %6.1f s  stopped at it=%d  Energy(1)=%.6e  -> %d KiB" % (name, time.time() - t, int(out["it_end"]), out["Energy"][0],
                                                                               os.path.getsize(path) // 1024), flush=True)
