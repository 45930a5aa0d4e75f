"""Coarse-to-fine driver around the QGMAP solver: the author's route to large displacements.

Mirrors legacy/optical_flow_ctf.m:21-36 (the pyramid loop) and legacy/gqmap_ctf.m (the single-Gaussian solver it calls).  As in the
reference, the pyramid glue (imresize, interp2 warp, fillmissing) runs on the HOST between solver calls -- it touches each
pixel once per scale -- and every ascent iteration runs in the CUDA iteration kernel through the C ABI (Solver).

What is the same as the reference: the scale schedule, warp accumulation `warp = imresize(warp,2)*2 + flow`, backward warp of
the first frame by linear interp2 with nearest fill of the out-of-frame pixels, and gqmap_ctf's solver constants
(legacy/gqmap_ctf.m:7,19-20,36,46-51: L=1, constant step 0.07, sigma stepped with step*0.3, sigma in [0.01,25], correlation
clamp 0.999, sigma init rand+3, no entropy term, clamp range = extrema of the scaled ground truth).
What is deliberately different (B200-first, documented in DESIGN.md): the data term samples the second frame with the
live solver's exact bicubic (gqmap_gpu_mixture.m:156-179) instead of a nearest lookup into a 64x bicubically upsampled copy
(legacy/gqmap_ctf.m:10,76: 4096x the image in memory, positions quantised to 1/64 px); the warp is resized to the exact size of
the next level (the reference's fixed factor 2 only works when every level size is even); AEPE per level is measured against
the ground truth resized to that level (the reference indexes the full-size array with the small level's ranges).
"""
import numpy as np

from .host import Solver


def _cubic(x):
    """MATLAB imresize's bicubic kernel (Keys, a = -0.5)."""
    ax = np.abs(x)
    return ((1.5 * ax ** 3 - 2.5 * ax ** 2 + 1) * (ax <= 1) +
            (-0.5 * ax ** 3 + 2.5 * ax ** 2 - 4 * ax + 2) * ((1 < ax) & (ax <= 2)))


def _contributions(in_len, out_len, scale, antialiasing=True):
    """imresize.m `contributions`: for every output sample the input indices (0-based) and normalised weights."""
    kw = 4.0
    if scale < 1 and antialiasing:
        h = lambda x: scale * _cubic(scale * x)
        kw = kw / scale
    else:
        h = _cubic
    x = np.arange(1, out_len + 1, dtype=np.float64)[:, None]
    u = x / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - kw / 2)
    P = int(np.ceil(kw)) + 2
    ind = left + np.arange(P)[None, :]
    w = h(u - ind)
    w = w / w.sum(axis=1, keepdims=True)
    aux = np.concatenate([np.arange(1, in_len + 1), np.arange(in_len, 0, -1)])
    ind = aux[np.mod(ind.astype(np.int64) - 1, aux.size)] - 1
    return w, ind


def imresize(A, scale=None, size=None):
    """B = imresize(A, scale) / imresize(A, [rows cols]) for double arrays: bicubic, antialiased when shrinking, dimensions
    resized in order of increasing scale (rows first on ties), symmetric boundary -- MATLAB's algorithm."""
    A = np.asarray(A, dtype=np.float64)
    M, N = A.shape[:2]
    if size is not None:
        oM, oN = int(size[0]), int(size[1])
        sc = (oM / M, oN / N)
    else:
        sc = (float(scale), float(scale))
        oM, oN = int(np.ceil(M * sc[0])), int(np.ceil(N * sc[1]))
    out = A
    for dim in sorted((0, 1), key=lambda d: sc[d]):
        w, ind = _contributions((M, N)[dim], (oM, oN)[dim], sc[dim])
        src = np.moveaxis(out, dim, 0)
        res = np.einsum("op,op...->o...", w, src[ind])
        out = np.moveaxis(res, 0, dim)
    return np.ascontiguousarray(out)


def interp2_linear(V, xq, yq):
    """Vq = interp2(V, xq, yq): bilinear, 1-based query coordinates, NaN outside the image (MATLAB defaults)."""
    V = np.asarray(V, dtype=np.float64)
    M, N = V.shape
    x, y = np.asarray(xq, dtype=np.float64) - 1.0, np.asarray(yq, dtype=np.float64) - 1.0
    ok = (x >= 0) & (x <= N - 1) & (y >= 0) & (y <= M - 1)
    xc, yc = np.clip(x, 0, N - 1), np.clip(y, 0, M - 1)
    x0 = np.minimum(np.floor(xc).astype(np.int64), N - 2) if N > 1 else np.zeros_like(xc, dtype=np.int64)
    y0 = np.minimum(np.floor(yc).astype(np.int64), M - 2) if M > 1 else np.zeros_like(yc, dtype=np.int64)
    s, t = xc - x0, yc - y0
    x1, y1 = np.minimum(x0 + 1, N - 1), np.minimum(y0 + 1, M - 1)
    out = (V[y0, x0] * (1 - s) + V[y0, x1] * s) * (1 - t) + (V[y1, x0] * (1 - s) + V[y1, x1] * s) * t
    return np.where(ok, out, np.nan)


def fillmissing_nearest(A, axis):
    """fillmissing(A,'nearest',dim): NaNs take the nearest non-NaN value along `axis` (the later one on ties); lines that are
    all NaN stay NaN."""
    A = np.array(A, dtype=np.float64, copy=True)
    B = np.moveaxis(A, axis, 0)
    n = B.shape[0]
    idx = np.arange(n).reshape((n,) + (1,) * (B.ndim - 1))
    good = ~np.isnan(B)
    prev = np.maximum.accumulate(np.where(good, idx, -1), axis=0)
    nxt = np.flip(np.minimum.accumulate(np.flip(np.where(good, idx, 2 * n), axis=0), axis=0), axis=0)
    dp, dn = np.where(prev >= 0, idx - prev, 4 * n), np.where(nxt < n, nxt - idx, 4 * n)
    src = np.where(dn <= dp, nxt, prev)
    have = (prev >= 0) | (nxt < n)
    filled = np.take_along_axis(B, np.clip(src, 0, n - 1), axis=0)
    B[...] = np.where(good, B, np.where(have, filled, np.nan))
    return A


def ctf_options(options, grdt):
    """The constants legacy/gqmap_ctf.m hard-codes, as options of the live solver."""
    o = dict(options)
    o.update(L=1, temperature=0.0, drate=1.0,
             minu=float(grdt[:, :, 0].min()), maxu=float(grdt[:, :, 0].max()),                  # gqmap_ctf.m:4
             minv=float(grdt[:, :, 1].min()), maxv=float(grdt[:, :, 1].max()),
             step0=0.07, step_tau=1e300, sigma_step_scale=0.3, sigma_min=0.01, sigma_max=25.0, corr_tor=0.999)   # :7,:36,:46-51
    return o


def gqmap_ctf(options, I1, I2, GRDT, seed=0, device=-1, aepe_target=None):
    """[mu, sigma, rou, AEPE, Energy] = gqmap_ctf(options, I1, I2, GRDT) (legacy/gqmap_ctf.m:1) on the CUDA solver.
    GRDT supplies the clamp range of the means (its extrema, gqmap_ctf.m:4).  The reference also evaluates AEPE against GRDT
    every iteration (:52), which would cost a state read-back per iteration: here AEPE is reported for the final beliefs only
    (entry n_done-1, NaN elsewhere), against aepe_target (default GRDT) when that has the level's size."""
    I1, I2 = np.asfortranarray(I1, dtype=np.float64), np.asfortranarray(I2, dtype=np.float64)
    M, N = I1.shape
    o = ctf_options(options, GRDT)
    o["device"] = device
    its = int(o["its"])
    rng = np.random.default_rng(seed)
    init = dict(w=np.zeros(1), muu=o["minu"] + rng.random((M, N, 1)) * (o["maxu"] - o["minu"]),         # gqmap_ctf.m:14-19
                muv=o["minv"] + rng.random((M, N, 1)) * (o["maxv"] - o["minv"]),
                sigmau=rng.random((M, N, 1)) + 3.0, sigmav=rng.random((M, N, 1)) + 3.0,
                pn=np.zeros((M, N, 1)), rou=np.zeros((M, N, 1, 2, 2)))
    with Solver(o, I1, I2) as s:
        s.set_state(init, T=0.0)
        r = s.step(its, its=its)
        st = s.get_state()
    mu = np.asfortranarray(np.concatenate([st["muu"], st["muv"]], axis=2))
    sigma = np.asfortranarray(np.concatenate([st["sigmau"], st["sigmav"]], axis=2))
    AEPE = np.full(its, np.nan)
    tgt = GRDT if aepe_target is None else aepe_target
    if tgt is not None and tgt.shape[:2] == (M, N):
        d = np.sqrt(((tgt[1:-1, 1:-1] - mu[1:-1, 1:-1]) ** 2).sum(axis=2))                              # gqmap_ctf.m:52
        AEPE[max(r["n_done"], 1) - 1] = d.mean()
    Energy = np.zeros(its)
    Energy[:r["n_done"]] = r["Energy"]
    return mu, sigma, st["rou"][:, :, 0], AEPE, Energy


def optical_flow_ctf(img_1, img_2, trueFlow, options, scales=(1 / 8, 1 / 4, 1 / 2, 1), seed=0, device=-1, verbose=False):
    """The pyramid loop of legacy/optical_flow_ctf.m:21-36.  img_1, img_2: grey frames (double); trueFlow: M x N x 2 with unknown
    flow already zeroed (flowToColor_mex's second output; required, as in the reference: it fixes the clamp range).  Returns (warp, levels): the accumulated full-resolution flow and a
    list of per-level dicts (scale, shape, aepe_before, aepe_after, iterations, ms)."""
    img_1, img_2 = np.asarray(img_1, dtype=np.float64), np.asarray(img_2, dtype=np.float64)
    M, N = img_1.shape
    warp = None
    levels = []
    for scale in scales:
        I1 = imresize(img_1, scale)                                             # :24-25
        I2 = imresize(img_2, scale)
        m, n = I1.shape
        if warp is None:
            warp = np.zeros((m, n, 2))                                          # :22 (zeros resized stay zeros)
        else:
            fy, fx = m / warp.shape[0], n / warp.shape[1]                       # :27 `imresize(warp,2).*2`, exact level sizes
            warp = imresize(warp, size=(m, n)) * np.array([fx, fy])[None, None, :]
        x, y = np.meshgrid(np.arange(1, n + 1, dtype=np.float64), np.arange(1, m + 1, dtype=np.float64))
        I1_w = interp2_linear(I1, x - warp[:, :, 0], y - warp[:, :, 1])         # :28-29
        I1_w = fillmissing_nearest(fillmissing_nearest(I1_w, 0), 1)             # :30
        gt = imresize(trueFlow, size=(m, n)) * scale
        resid = gt - warp                                                       # what this level still has to find
        mu, sigma, rou, AEPE, Energy = gqmap_ctf(options, I1_w, I2, trueFlow * scale, seed=seed, device=device,   # :31 `trueFlow.*scale`
                                                 aepe_target=resid)
        lv = dict(scale=scale, shape=(m, n), iterations=int(np.count_nonzero(Energy)))
        e = lambda f: float(np.sqrt(((gt - f)[1:-1, 1:-1] ** 2).sum(axis=2)).mean()) / scale          # in full-resolution pixels
        lv["aepe_before"], lv["aepe_after"] = e(warp), e(warp + mu)
        warp = warp + mu                                                        # :32
        levels.append(lv)
        if verbose:
            print(lv, flush=True)
    return warp, levels
