"""Frame / flow I/O helpers used by the drivers, the tests and the bench (host glue, not the hot path).

readFlowFile mirrors readFlowFile.m:33-81; rgb2gray mirrors MATLAB's rgb2gray on uint8 (optical_flow.m:10-11);
synthetic_pair builds the synthetic workloads named in BASELINE.json.
"""
import numpy as np

TAG_FLOAT = 202021.25      # readFlowFile.m:33


def readFlowFile(filename):
    """img = readFlowFile(filename): Middlebury .flo -> H x W x 2 float64 (u,v)."""
    if not filename:
        raise ValueError("readFlowFile: empty filename")
    if not filename.endswith(".flo"):
        raise ValueError("readFlowFile: filename %s should have extension '.flo'" % filename)
    with open(filename, "rb") as f:
        tag = np.fromfile(f, np.float32, 1)
        wh = np.fromfile(f, np.int32, 2)
        if tag.size != 1 or tag[0] != np.float32(TAG_FLOAT):
            raise ValueError("readFlowFile(%s): wrong tag (possibly due to big-endian machine?)" % filename)
        width, height = int(wh[0]), int(wh[1])
        if width < 1 or width > 99999:
            raise ValueError("readFlowFile(%s): illegal width %d" % (filename, width))
        if height < 1 or height > 99999:
            raise ValueError("readFlowFile(%s): illegal height %d" % (filename, height))
        tmp = np.fromfile(f, np.float32)
    tmp = tmp.reshape(height, width * 2).astype(np.float64)
    return np.asfortranarray(np.stack([tmp[:, 0::2], tmp[:, 1::2]], axis=2))


def writeFlowFile(img, filename):
    """writeFlowFile(img, filename): H x W x 2 flow -> Middlebury .flo (legacy/writeFlowFile.m:27-76)."""
    if not filename:
        raise ValueError("writeFlowFile: empty filename")
    if not filename.endswith(".flo"):
        raise ValueError("writeFlowFile: filename %s should have extension '.flo'" % filename)
    img = np.asarray(img)
    if img.ndim != 3 or img.shape[2] != 2:
        raise ValueError("writeFlowFile: image must have two bands")
    height, width, _ = img.shape
    with open(filename, "wb") as f:
        f.write(b"PIEH")
        np.array([width, height], np.int32).tofile(f)
        tmp = np.empty((height, width * 2), np.float32)
        tmp[:, 0::2] = img[:, :, 0]
        tmp[:, 1::2] = img[:, :, 1]
        tmp.tofile(f)


def save_results(path, options, mu, sigma, alpha, AEPE, Energy, logP):
    """save([options.dir '/' name '.mat'],'options','AEPE','mu','sigma','alpha','Energy','logP') -- optical_flow.m:28."""
    from scipy.io import savemat
    opts = {k: (np.asarray(v) if not isinstance(v, str) else v) for k, v in dict(options).items() if v is not None and k != "init"}
    savemat(path, dict(options=opts, AEPE=AEPE, mu=mu, sigma=sigma, alpha=alpha, Energy=Energy, logP=logP), do_compression=True)


def rgb2gray(rgb):
    """MATLAB rgb2gray on uint8: 0.298936021293775 R + 0.587043074451121 G + 0.114020904255103 B, rounded to uint8."""
    rgb = np.asarray(rgb)
    if rgb.ndim == 2:
        return rgb.astype(np.uint8)
    g = rgb[..., :3].astype(np.float64) @ np.array([0.298936021293775, 0.587043074451121, 0.114020904255103])
    return np.clip(np.floor(g + 0.5), 0, 255).astype(np.uint8)


# the Middlebury training sequences shipped by the reference (rows, cols): SURVEY.md section 2
middlebury_shapes = {
    "RubberWhale": (388, 584), "Dimetrodon": (388, 584), "Hydrangea": (388, 584), "Venus": (380, 420),
    "Grove2": (480, 640), "Grove3": (480, 640), "Urban2": (480, 640), "Urban3": (480, 640),
}


def _gauss_blur(a, sigma):
    r = int(np.ceil(4 * sigma))
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    p = np.pad(a, r, mode="reflect")
    p = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 0, p)
    return np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 1, p)


def _bicubic(V, yq, xq):
    """Keys(-0.5) bicubic sampling of V at 0-based (yq,xq) with replicate padding (workload generation only)."""
    M, N = V.shape
    P = np.pad(V, 2, mode="edge")
    yq = np.clip(yq, 0, M - 1)
    xq = np.clip(xq, 0, N - 1)
    iy = np.minimum(np.floor(yq).astype(np.int64), M - 2)
    ix = np.minimum(np.floor(xq).astype(np.int64), N - 2)
    t, s = yq - iy, xq - ix

    def w(s):
        return (((2 - s) * s - 1) * s / 2, ((3 * s - 5) * s * s + 2) / 2, (((4 - 3 * s) * s + 1) * s) / 2, ((s - 1) * s * s) / 2)
    ws, wt = w(s), w(t)
    out = 0.0
    for r in range(4):
        for c in range(4):
            out = out + P[iy + r + 1, ix + c + 1] * wt[r] * ws[c]
    return out


def synthetic_pair(M, N, seed=1234, flow_scale=1.0, grey_levels=False):
    """Synthetic frame pair of SURVEY.md section 8d: I1 = Gaussian-blurred (sigma 1.5) uniform noise scaled to [0,255];
    ground-truth flow u = 3 sin(2 pi row/M) + 1 (horizontal), v = 2 cos(2 pi col/N) (vertical); I2 = I1 warped so that
    I2(row+v, col+u) = I1(row,col) (bicubic resampling of I1 at the inverse flow, found by fixed-point iteration).
    grey_levels=True rounds both frames to integer grey levels 0..255, which is what the reference's drivers feed the solver
    (double(rgb2gray(imresize(imread(...)))), optical_flow.m:8-11) -- the ground truth is then exact up to that quantisation.
    Returns I1, I2 (M x N float64, 0..255), flow (M x N x 2), and (minu,maxu,minv,maxv)."""
    rng = np.random.default_rng(seed)
    a = _gauss_blur(rng.random((M, N)), 1.5)
    I1 = (a - a.min()) / (a.max() - a.min()) * 255.0
    rows = np.arange(M, dtype=np.float64).reshape(M, 1)
    cols = np.arange(N, dtype=np.float64).reshape(1, N)
    fu = lambda r, c: flow_scale * (3.0 * np.sin(2 * np.pi * r / M) + 1.0) + 0.0 * c      # flow_scale > 1: large displacements
    fv = lambda r, c: flow_scale * (2.0 * np.cos(2 * np.pi * c / N)) + 0.0 * r            # (coarse-to-fine workloads)
    u, v = fu(rows, cols), fv(rows, cols)
    # I2(q) = I1(p) with p + flow(p) = q: invert the (smooth, contractive) flow by fixed-point iteration
    pr, pc = rows - v, cols - u
    for _ in range(30):
        pr, pc = rows - fv(pr, pc), cols - fu(pr, pc)
    I2 = _bicubic(I1, pr, pc)
    if grey_levels:
        I1, I2 = np.clip(np.round(I1), 0.0, 255.0), np.clip(np.round(I2), 0.0, 255.0)
    flow = np.asfortranarray(np.stack([u, v], axis=2))
    return np.asfortranarray(I1), np.asfortranarray(I2), flow, (float(u.min()), float(u.max()), float(v.min()), float(v.max()))
