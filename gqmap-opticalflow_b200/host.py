"""Host-side mirror of the reference's MATLAB interface for the QGMAP hot path, over the C ABI of libqgmap.so.

MATLAB is the reference's host language but is not installed where this code is built and tested, so the same
operator boundary is mirrored here one-to-one (same names, argument meaning, returned arrays and shapes); the MATLAB
wrappers in matlab/ and the MEX gateways in mex/ bind the very same C symbols.

  [mu,sigma,alpha,AEPE,Energy,logP] = gqmap_gpu_mixture(options,I1,I2)          gqmap_gpu_mixture.m:1
  [mu,sigma,alpha,AEPE,Energy,logP] = gqmap_gpuSuper_mix_entropy(options,I1,I2) gqmap_gpuSuper_mix_entropy.m:1
  map = get_map_mex(alf, mu_u, sig_u, mu_v, sig_v)                              gqmap_gpu_mixture.m:57
  [img,flo,minu,maxu,minv,maxv,idxUnknown] = flowToColor_mex(flow[,maxFlow])    optical_flow.m:12-13
  [x,w] = GaussHermite_2(n)                                                     GaussHermite_2.m:1
  x = projsplx(y)                                                               projsplx.m:1

`options` is a dict (or any object with attributes) carrying the reference's fields (gqmap_gpu_mixture.m:3-6):
trueFlow, unknownIdx, its, K, L, temperature, drate, epsn, lambdad, lambdas, minu, maxu, minv, maxv [, dir].
New OPTIONAL fields only: init (dict of the 7 state arrays muu,muv,sigmau,sigmav,pn,rou,w), seed, alpha_mode
('softmax'|'projsplx'), device, devices (list of CUDA ordinals: the frame pair is split into one row band per entry), log_every.  options.dir, when given, receives <it>.png at every monitored iteration
(:59-62); the directory must exist (the drivers mkdir it, optical_flow.m:25).  Everything runs on the GPU; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import QgmapConfig, QgmapError, check, dptr, f64, lib, u8ptr

STATE_FIELDS = ("muu", "muv", "sigmau", "sigmav", "pn", "rou", "w")


def _opt(options, name, default=None, required=False):
    if isinstance(options, dict):
        if name in options:
            return options[name]
    elif hasattr(options, name):
        return getattr(options, name)
    if required:
        raise KeyError("options.%s is required (gqmap_gpu_mixture.m:3-6)" % name)
    return default


def make_config(options, variant):
    """Build qgmap_config from the reference's options struct (gqmap_gpu_mixture.m:3-6)."""
    cfg = QgmapConfig()
    check(lib.qgmap_config_defaults(C.byref(cfg), variant))
    cfg.K = int(_opt(options, "K", required=True))
    cfg.L = int(_opt(options, "L", required=True))
    cfg.temperature = float(_opt(options, "temperature", required=True))
    cfg.drate = float(_opt(options, "drate", required=True))
    cfg.epsn = float(_opt(options, "epsn", required=True))
    cfg.lambdad = float(_opt(options, "lambdad", required=True))
    cfg.lambdas = float(_opt(options, "lambdas", required=True))
    for k in ("minu", "maxu", "minv", "maxv"):
        setattr(cfg, k, float(_opt(options, k, required=True)))
    mode = _opt(options, "alpha_mode", "softmax")
    if mode not in ("softmax", "projsplx"):
        raise ValueError("options.alpha_mode must be 'softmax' or 'projsplx'")
    cfg.alpha_mode = _lib.ALPHA_PROJSPLX if mode == "projsplx" else _lib.ALPHA_SOFTMAX
    cfg.device = int(_opt(options, "device", -1))
    cfg.log_every = int(_opt(options, "log_every", 300))
    for k in ("sigma_min", "sigma_max", "corr_tor", "step0", "step_tau", "alpha_scale", "T_floor", "tor", "sigma_step_scale"):
        v = _opt(options, k)
        if v is not None:
            setattr(cfg, k, float(v))
    for k in ("alpha_start", "anneal_every", "row_begin", "row_end", "strip_rows"):
        v = _opt(options, k)
        if v is not None:
            setattr(cfg, k, int(v))
    return cfg


def _check_images(I1, I2, variant):
    I1, I2 = f64(I1), f64(I2)
    if I1.ndim != 2 or I1.shape != I2.shape:
        raise ValueError("I1 and I2 must be 2-D arrays of equal size")
    if variant == _lib.VARIANT_SUPER and (I1.shape[0] % 4 or I1.shape[1] % 4):
        raise ValueError("gqmap_gpuSuper_mix_entropy needs image sides divisible by 4 (M=Mo/4, N=No/4)")
    return I1, I2


def _solve(options, I1, I2, variant):
    I1, I2 = _check_images(I1, I2, variant)
    Mo, No = I1.shape
    cfg = make_config(options, variant)
    its = int(_opt(options, "its", required=True))
    M, N = (Mo // 4, No // 4) if variant == _lib.VARIANT_SUPER else (Mo, No)
    L = cfg.L
    tflow = _opt(options, "trueFlow")
    unk = _opt(options, "unknownIdx")
    if tflow is not None:
        tflow = f64(tflow)
        if tflow.shape != (Mo, No, 2):
            raise ValueError("options.trueFlow must be Mo x No x 2")
    if unk is not None:
        unk = np.asfortranarray(np.asarray(unk, dtype=np.uint8))
        if unk.shape != (Mo, No):
            raise ValueError("options.unknownIdx must be Mo x No")
    init = _opt(options, "init")
    init_arr = None
    keep = []
    if init is not None:
        shapes = {"muu": (M, N, L), "muv": (M, N, L), "sigmau": (M, N, L), "sigmav": (M, N, L), "pn": (M, N, L),
                  "rou": (M, N, L, 2, 2), "w": (L,)}
        init_arr = (C.POINTER(C.c_double) * 7)()
        for i, name in enumerate(STATE_FIELDS):
            a = f64(np.asarray(init[name], dtype=np.float64).reshape(shapes[name], order="F"))
            keep.append(a)
            init_arr[i] = dptr(a)
    mu = np.zeros((M, N, L, 2), order="F")
    sigma = np.zeros((M, N, L, 2), order="F")
    alpha = np.zeros(L)
    AEPE = np.zeros(its)
    Energy = np.zeros(its)
    logP = np.zeros(its)
    done = C.c_int(0)
    out_dir = _opt(options, "dir")                  # gqmap_gpu_mixture.m:62: [options.dir '/' num2str(it) '.png']
    check(lib.qgmap_solve_set_dump_dir(None if not out_dir else str(out_dir).encode()))
    devices = _opt(options, "devices")             # new optional field: one row band of the frame pair per listed CUDA device
    try:
        if devices is not None and len(devices) > 1:
            dev = (C.c_int * len(devices))(*[int(d) for d in devices])
            check(lib.qgmap_group_solve(C.byref(cfg), dptr(I1), dptr(I2), Mo, No, its, len(devices), dev, init_arr,
                                        C.c_uint64(int(_opt(options, "seed", 0))), dptr(tflow), u8ptr(unk),
                                        dptr(mu), dptr(sigma), dptr(alpha), dptr(AEPE), dptr(Energy), dptr(logP), C.byref(done)))
        else:
            check(lib.qgmap_solve(C.byref(cfg), dptr(I1), dptr(I2), Mo, No, its, init_arr,
                                  C.c_uint64(int(_opt(options, "seed", 0))), dptr(tflow), u8ptr(unk),
                                  dptr(mu), dptr(sigma), dptr(alpha), dptr(AEPE), dptr(Energy), dptr(logP), C.byref(done)))
    finally:
        lib.qgmap_solve_set_dump_dir(None)
    return mu, sigma, alpha.reshape(1, 1, L), AEPE.reshape(its, 1), Energy.reshape(its, 1), logP.reshape(its, 1)


def gqmap_gpu_mixture(options, I1, I2):
    """Full-resolution QGMAP solver; same signature and outputs as gqmap_gpu_mixture.m:1,183-188."""
    return _solve(options, I1, I2, _lib.VARIANT_FULL)


def gqmap_gpuSuper_mix_entropy(options, I1, I2):
    """Super-pixel (4x4 block) QGMAP solver; same signature and outputs as gqmap_gpuSuper_mix_entropy.m:1,199-204."""
    return _solve(options, I1, I2, _lib.VARIANT_SUPER)


def last_solve_stats():
    """(kernel launches, device ms of the iteration kernels) of the last gqmap_* call on this thread."""
    n = C.c_longlong(0)
    ms = C.c_float(0)
    check(lib.qgmap_last_solve_stats(C.byref(n), C.byref(ms)))
    return n.value, ms.value


def get_map_mex(alf, mu_u, sig_u, mu_v, sig_v, device=-1):
    """map = get_map_mex(alf, mu_u, sig_u, mu_v, sig_v): per-pixel mixture mode of each flow layer (M x N x 2).
    Errors mirror the MEX's: five inputs of matching size are required."""
    arrs = [f64(a) for a in (mu_u, sig_u, mu_v, sig_v)]
    arrs = [a.reshape(a.shape + (1,), order="F") if a.ndim == 2 else a for a in arrs]
    if any(a.ndim != 3 or a.shape != arrs[0].shape for a in arrs):
        raise ValueError("get_map_mex: mu_u, sig_u, mu_v, sig_v must all be M x N x L")
    M, N, L = arrs[0].shape
    alf = np.ascontiguousarray(np.asarray(alf, dtype=np.float64).ravel())
    if alf.size != L:
        raise ValueError("get_map_mex: numel(alf) must equal size(mu_u,3)")
    out = np.zeros((M, N, 2), order="F")
    check(lib.qgmap_find_map(dptr(alf), *(dptr(a) for a in arrs), M, N, L, dptr(out), device))
    return out


def flowToColor_mex(flow, maxFlow=None):
    """[img,flo,minu,maxu,minv,maxv,idxUnknown] = flowToColor_mex(flow[,maxFlow]) (legacy/flowToColor.m:1)."""
    flow = f64(flow)
    if flow.ndim != 3 or flow.shape[2] != 2:
        raise ValueError("flowToColor: image must have two bands")          # legacy/flowToColor.m:41-43
    M, N, _ = flow.shape
    img = np.zeros((M, N, 3), dtype=np.uint8, order="F")
    flo = np.zeros((M, N, 2), order="F")
    stats = np.zeros(4)
    unk = np.zeros((M, N), dtype=np.uint8, order="F")
    check(lib.qgmap_flow_to_color(dptr(flow), M, N, -1.0 if maxFlow is None else float(maxFlow),
                                  u8ptr(img), dptr(flo), dptr(stats), u8ptr(unk)))
    return img, flo, stats[0], stats[1], stats[2], stats[3], unk.astype(bool)


def imwrite(img, filename):
    """imwrite(flc, filename) for the M x N x 3 uint8 images flowToColor_mex returns (gqmap_gpu_mixture.m:62)."""
    img = np.asfortranarray(np.asarray(img, dtype=np.uint8))
    if img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("imwrite: M x N x 3 uint8 image expected")
    check(lib.qgmap_write_png(str(filename).encode(), u8ptr(img), img.shape[0], img.shape[1]))


def GaussHermite_2(n):
    x = np.zeros(n)
    w = np.zeros(n)
    check(lib.qgmap_gauss_hermite(int(n), dptr(x), dptr(w)))
    return x.reshape(n, 1), w.reshape(n, 1)


def projsplx(y):
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
    x = np.zeros_like(y)
    check(lib.qgmap_projsplx(dptr(y), y.size, dptr(x)))
    return x


class Solver:
    """Stateful handle == the gqmap_mex('create'|'set_state'|'get_state'|'step'|'map'|'logp'|'destroy') interface."""

    def __init__(self, options, I1, I2, variant="full"):
        self.variant = _lib.VARIANT_SUPER if variant in ("super", _lib.VARIANT_SUPER) else _lib.VARIANT_FULL
        I1, I2 = _check_images(I1, I2, self.variant)
        self.cfg = make_config(options, self.variant)
        self.Mo, self.No = I1.shape
        self._h = C.c_void_p(None)
        check(lib.qgmap_create(C.byref(self.cfg), dptr(I1), dptr(I2), self.Mo, self.No, C.byref(self._h)))
        M, N, L = C.c_int(), C.c_int(), C.c_int()
        check(lib.qgmap_dims(self._h, C.byref(M), C.byref(N), C.byref(L)))
        self.M, self.N, self.L = M.value, N.value, L.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.qgmap_destroy(self._h)
            self._h = C.c_void_p(None)

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _shape(self, name):
        return {"rou": (self.M, self.N, self.L, 2, 2), "w": (self.L,)}.get(name, (self.M, self.N, self.L))

    def set_state(self, state, T=None, it=1, alpha=None):
        """State arrays as the reference holds them (double) -- or all six belief arrays as float32 (MATLAB `single`): the device
        keeps fp32, so that boundary format is lossless and moves half the bytes (qgmap_set_state_f32)."""
        al = None if alpha is None else np.ascontiguousarray(np.asarray(alpha, dtype=np.float64).ravel())
        Tv = self.cfg.temperature if T is None else float(T)
        beliefs = [n for n in STATE_FIELDS if n != "w"]
        if all(getattr(state[n], "dtype", None) == np.float32 for n in beliefs):
            arrs = [np.asfortranarray(np.asarray(state[n], dtype=np.float32).reshape(self._shape(n), order="F")) for n in beliefs]
            w = np.ascontiguousarray(np.asarray(state["w"], dtype=np.float64).ravel())
            check(lib.qgmap_set_state_f32(self._h, *(C.c_void_p(a.ctypes.data) for a in arrs), dptr(w), dptr(al), Tv, int(it)), self._h)
            return
        arrs = [f64(np.asarray(state[n], dtype=np.float64).reshape(self._shape(n), order="F")) for n in STATE_FIELDS]
        check(lib.qgmap_set_state(self._h, *(dptr(a) for a in arrs), dptr(al), Tv, int(it)), self._h)

    def init_state(self, seed=0):
        check(lib.qgmap_init_state(self._h, C.c_uint64(int(seed))), self._h)

    def get_state(self, dtype=np.float64):
        """dtype=np.float32 returns the belief arrays as the device holds them (qgmap_get_state_f32): lossless, half the bytes."""
        alpha = np.zeros(self.L)
        T = C.c_double(0)
        it = C.c_int(0)
        if np.dtype(dtype) == np.float32:
            out = {n: np.zeros(self._shape(n), dtype=np.float32 if n != "w" else np.float64, order="F") for n in STATE_FIELDS}
            check(lib.qgmap_get_state_f32(self._h, *(C.c_void_p(out[n].ctypes.data) for n in STATE_FIELDS if n != "w"), dptr(out["w"]),
                                          dptr(alpha), C.byref(T), C.byref(it)), self._h)
        else:
            out = {n: np.zeros(self._shape(n), order="F") for n in STATE_FIELDS}
            check(lib.qgmap_get_state(self._h, *(dptr(out[n]) for n in STATE_FIELDS), dptr(alpha), C.byref(T), C.byref(it)), self._h)
        out["alpha"], out["T"], out["it"] = alpha, T.value, it.value
        return out

    def step(self, n, its=2 ** 30):
        """Run up to n iterations; returns dict(Energy, ptdmu, ptdsigma (each n_done long), n_done, stopped, ms)."""
        E, dm, ds = np.zeros(max(n, 1)), np.zeros(max(n, 1)), np.zeros(max(n, 1))
        done, stopped = C.c_int(0), C.c_int(0)
        check(lib.qgmap_step(self._h, int(n), int(its), dptr(E), dptr(dm), dptr(ds), C.byref(done), C.byref(stopped)), self._h)
        ms = C.c_float(0)
        lib.qgmap_last_step_ms(self._h, C.byref(ms))
        nl = C.c_longlong(0)
        lib.qgmap_last_launches(self._h, C.byref(nl))
        k = done.value
        return dict(Energy=E[:k], ptdmu=dm[:k], ptdsigma=ds[:k], n_done=k, stopped=bool(stopped.value), ms=ms.value,
                    launches=nl.value)

    def map(self):
        out = np.zeros((self.M, self.N, 2), order="F")
        check(lib.qgmap_get_map(self._h, dptr(out)), self._h)
        return out

    def logp(self, uv):
        uv = f64(uv)
        lp = C.c_double(0)
        check(lib.qgmap_logp(self._h, dptr(uv), C.byref(lp)), self._h)
        return lp.value

    def aepe(self, uv, tflow, unknown=None):
        uv, tflow = f64(uv), f64(tflow)
        unk = None if unknown is None else np.asfortranarray(np.asarray(unknown, dtype=np.uint8))
        v = C.c_double(0)
        check(lib.qgmap_aepe(self._h, dptr(uv), dptr(tflow), u8ptr(unk), C.byref(v)), self._h)
        return v.value

    def set_truth(self, tflow, unknown=None):
        """options.trueFlow / options.unknownIdx for monitor_partial (uploaded once)."""
        tflow = f64(tflow)
        if tflow.shape != (self.Mo, self.No, 2):
            raise ValueError("trueFlow must be Mo x No x 2")
        unk = None if unknown is None else np.asfortranarray(np.asarray(unknown, dtype=np.uint8))
        check(lib.qgmap_set_truth(self._h, dptr(tflow), u8ptr(unk)), self._h)

    def monitor_partial(self, aepe=True):
        """The monitoring block (gqmap_gpu_mixture.m:52-67) for the rows this handle owns, on its device: (share of logP, share of
        the AEPE numerator).  A row-band run adds the shares of all bands and divides the AEPE sum by (Mo-2b)(No-2b)."""
        lp, ae = C.c_double(0), C.c_double(0)
        check(lib.qgmap_monitor_partial(self._h, C.byref(lp), C.byref(ae) if aepe else None), self._h)
        return lp.value, ae.value

    def debug_gradients(self):
        names = ("G_muu", "G_muv", "G_sigu", "G_sigv", "dpn", "drou", "e_px", "da_px")
        out = {n: np.zeros((self.M, self.N, self.L, 2, 2) if n == "drou" else (self.M, self.N, self.L), order="F") for n in names}
        check(lib.qgmap_debug_gradients(self._h, *(dptr(out[n]) for n in names)), self._h)
        return out


def batch_step(solvers, n, its=2 ** 30):
    """Run n iterations on every Solver of a same-device batch concurrently; returns (device_ms, kernel launches)."""
    arr = (C.c_void_p * len(solvers))(*[s._h for s in solvers])
    ms = C.c_float(0)
    nl = C.c_longlong(0)
    check(lib.qgmap_batch_step(arr, len(solvers), int(n), int(its), C.byref(ms), C.byref(nl)))
    return ms.value, nl.value


def fp32_peak(device=-1):
    """Measured FP32 FMA throughput of the device in TFLOP/s (roofline denominator)."""
    v = C.c_double(0)
    check(lib.qgmap_fp32_peak(int(device), C.byref(v)))
    return v.value


class BandGroup:
    """One frame pair split into row bands, all bands driven by this process: qgmap_group_* of the C ABI.
    devices: list of CUDA ordinals, one per band (None: every band on the current device)."""

    def __init__(self, options, I1, I2, nbands, devices=None, variant="full"):
        self.variant = _lib.VARIANT_SUPER if variant in ("super", _lib.VARIANT_SUPER) else _lib.VARIANT_FULL
        I1, I2 = _check_images(I1, I2, self.variant)
        self.cfg = make_config(options, self.variant)
        self._g = C.c_void_p(None)
        dev = None
        if devices is not None:
            if len(devices) != nbands:
                raise ValueError("devices must list one CUDA ordinal per band")
            dev = (C.c_int * nbands)(*devices)
        check(lib.qgmap_group_create(C.byref(self.cfg), dptr(I1), dptr(I2), I1.shape[0], I1.shape[1], int(nbands), dev, C.byref(self._g)))
        M, N, L, nb = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib.qgmap_group_dims(self._g, C.byref(M), C.byref(N), C.byref(L), C.byref(nb)))
        self.M, self.N, self.L, self.nbands = M.value, N.value, L.value, nb.value

    def close(self):
        if getattr(self, "_g", None) is not None and self._g.value:
            lib.qgmap_group_destroy(self._g)
            self._g = C.c_void_p(None)

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    _shape = Solver._shape

    def set_state(self, state, T=None, it=1, alpha=None):
        arrs = [f64(np.asarray(state[n], dtype=np.float64).reshape(self._shape(n), order="F")) for n in STATE_FIELDS]
        al = None if alpha is None else np.ascontiguousarray(np.asarray(alpha, dtype=np.float64).ravel())
        check(lib.qgmap_group_set_state(self._g, *(dptr(a) for a in arrs), dptr(al),
                                        self.cfg.temperature if T is None else float(T), int(it)))

    def init_state(self, seed=0):
        check(lib.qgmap_group_init_state(self._g, C.c_uint64(int(seed))))

    def get_state(self):
        out = {n: np.zeros(self._shape(n), order="F") for n in STATE_FIELDS}
        alpha = np.zeros(self.L)
        T = C.c_double(0)
        it = C.c_int(0)
        check(lib.qgmap_group_get_state(self._g, *(dptr(out[n]) for n in STATE_FIELDS), dptr(alpha), C.byref(T), C.byref(it)))
        out["alpha"], out["T"], out["it"] = alpha, T.value, it.value
        return out

    def step(self, n, its=2 ** 30):
        E, dm, ds = np.zeros(max(n, 1)), np.zeros(max(n, 1)), np.zeros(max(n, 1))
        done, stopped = C.c_int(0), C.c_int(0)
        check(lib.qgmap_group_step(self._g, int(n), int(its), dptr(E), dptr(dm), dptr(ds), C.byref(done), C.byref(stopped)))
        ms = C.c_float(0)
        lib.qgmap_group_last_step_ms(self._g, C.byref(ms))
        k = done.value
        return dict(Energy=E[:k], ptdmu=dm[:k], ptdsigma=ds[:k], n_done=k, stopped=bool(stopped.value), ms=ms.value)
