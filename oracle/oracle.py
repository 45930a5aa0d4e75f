"""ctypes front-end of the CPU fp64 oracle (oracle/qgmap_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  PINNED against the reference, executed (see qgmap_oracle.h): find_map and flow_to_color against the shipped
.mexw64 binaries (oracle/refbin/), the iteration loop against the reference's own .m files run by oracle/mlab/minimat.py.

All arrays are MATLAB-shaped NumPy fp64 arrays in Fortran (column-major) order, e.g. muu is (M,N,L),
rou is (M,N,L,2,2) -- exactly the shapes gqmap_gpu_mixture.m:18-24 allocates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqgmap_oracle.so")


def build(force=False):
    """Compile the C restatement with the committed Makefile (gcc, -ffp-contract=off, OpenMP)."""
    src = os.path.join(_HERE, "qgmap_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(src[:-1] + "h"))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libqgmap_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class QoConfig(C.Structure):
    _fields_ = [
        ("Mo", C.c_int), ("No", C.c_int), ("M", C.c_int), ("N", C.c_int), ("L", C.c_int), ("K", C.c_int),
        ("super", C.c_int),
        ("lambdad", C.c_double), ("lambdas", C.c_double), ("epsn", C.c_double),
        ("minu", C.c_double), ("maxu", C.c_double), ("minv", C.c_double), ("maxv", C.c_double),
        ("sigma_min", C.c_double), ("sigma_max", C.c_double), ("corr_tor", C.c_double),
        ("step0", C.c_double), ("step_tau", C.c_double),
        ("alpha_start", C.c_int), ("alpha_scale", C.c_double), ("alpha_mode", C.c_int),
        ("anneal_every", C.c_int), ("drate", C.c_double), ("T_floor", C.c_double), ("tor", C.c_double),
        ("sigma_step_scale", C.c_double), ("guard_a0", C.c_int), ("nthreads", C.c_int),
    ]


_DP = C.POINTER(C.c_double)


class QoState(C.Structure):
    _fields_ = [("muu", _DP), ("muv", _DP), ("sigu", _DP), ("sigv", _DP), ("pn", _DP), ("rou", _DP),
                ("w", _DP), ("alpha", _DP), ("T", C.c_double)]


class QoGrads(C.Structure):
    _fields_ = [(n, _DP) for n in ("dan", "dmuu", "dmuv", "dsigmau", "dsigmav", "dpn", "nEnergy",
                                   "dae", "dmu1", "dmu2", "dsigma1", "dsigma2", "drou", "eEnergy")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.qo_node_pot.restype = C.c_double
        _lib.qo_edge_pot.restype = C.c_double
        _lib.qo_profile_logp.restype = C.c_double
        _lib.qo_aepe.restype = C.c_double
        _lib.qo_fminbnd_mixture.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(_DP)


def _f64(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def make_config(Mo, No, L, K, *, super=False, lambdad=1.0, lambdas=5.0, epsn=1e-6,
                minu=-1.0, maxu=1.0, minv=-1.0, maxv=1.0, alpha_mode=0, nthreads=0, drate=0.5, **over):
    """Defaults are the constants hard-wired in gqmap_gpu_mixture.m (full) / gqmap_gpuSuper_mix_entropy.m (super)."""
    c = QoConfig()
    c.Mo, c.No = Mo, No
    c.M, c.N = (Mo // 4, No // 4) if super else (Mo, No)
    c.L, c.K, c.super = L, K, int(super)
    c.lambdad, c.lambdas, c.epsn = lambdad, lambdas, epsn
    c.minu, c.maxu, c.minv, c.maxv = minu, maxu, minv, maxv
    c.sigma_min, c.sigma_max = 0.01, (25.0 if super else 23.0)
    c.corr_tor = 1 - 1e-5
    c.step0, c.step_tau = (0.001, 4000.0) if super else (0.1, 8000.0)
    c.alpha_start, c.alpha_scale, c.alpha_mode = 500, 1e-7, alpha_mode
    c.anneal_every = 500 if super else 0
    c.drate, c.T_floor, c.tor = drate, 0.001, 1e-4
    c.sigma_step_scale = 1.0
    c.guard_a0 = 0 if super else 1
    c.nthreads = nthreads
    for k, v in over.items():
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    return c


class State:
    """The solver state of gqmap_gpu_mixture.m:18-24 as Fortran-ordered fp64 arrays."""
    FIELDS = ("muu", "muv", "sigu", "sigv", "pn", "rou", "w", "alpha")

    def __init__(self, muu, muv, sigu, sigv, pn, rou, w, alpha=None, T=0.0):
        self.muu, self.muv, self.sigu, self.sigv, self.pn, self.rou = map(_f64, (muu, muv, sigu, sigv, pn, rou))
        self.w = np.ascontiguousarray(np.asarray(w, dtype=np.float64).ravel())
        if alpha is None:
            alpha = np.exp(self.w) / np.sum(np.exp(self.w))       # gqmap_gpu_mixture.m:18
        self.alpha = np.ascontiguousarray(np.asarray(alpha, dtype=np.float64).ravel())
        self.T = float(T)

    def copy(self):
        return State(*(getattr(self, f).copy() for f in self.FIELDS), T=self.T)

    def c_struct(self):
        s = QoState()
        for f in self.FIELDS:
            setattr(s, f, _p(getattr(self, f)))
        s.T = self.T
        return s


def init_state(cfg, seed, T=0.0):
    """gqmap_gpu_mixture.m:18-24 with NumPy's default_rng(seed) standing in for MATLAB's rand stream."""
    rng = np.random.default_rng(seed)
    M, N, L = cfg.M, cfg.N, cfg.L
    w = rng.random(L)
    muu = cfg.minu + rng.random((M, N, L)) * (cfg.maxu - cfg.minu)
    muv = cfg.minv + rng.random((M, N, L)) * (cfg.maxv - cfg.minv)
    sigu = rng.random((M, N, L)) + (cfg.maxu - cfg.minu)
    sigv = rng.random((M, N, L)) + (cfg.maxv - cfg.minv)
    pn = np.zeros((M, N, L))
    rou = np.zeros((M, N, L, 2, 2))
    return State(muu, muv, sigu, sigv, pn, rou, w, T=T)


def gauss_hermite(n):
    x = np.zeros(n)
    w = np.zeros(n)
    rc = lib().qo_gauss_hermite(C.c_int(n), _p(x), _p(w))
    if rc != 0:
        raise RuntimeError("qo_gauss_hermite failed rc=%d" % rc)
    return x, w


def get_vv(V):
    V = _f64(V)
    M, N = V.shape
    VV = np.zeros((M + 2, N + 2), order="F")
    lib().qo_get_vv(_p(V), C.c_int(M), C.c_int(N), _p(VV))
    return VV


def node_pot(cfg, I1, VV, x1, x2, i, j):
    return lib().qo_node_pot(C.byref(cfg), _p(I1), _p(VV), C.c_double(x1), C.c_double(x2), C.c_int(i), C.c_int(j))


def edge_pot(cfg, x1, x2):
    return lib().qo_edge_pot(C.byref(cfg), C.c_double(x1), C.c_double(x2))


NODE_NAMES = ("dan", "dmuu", "dmuv", "dsigmau", "dsigmav", "dpn", "nEnergy")
EDGE_NAMES = ("dae", "dmu1", "dmu2", "dsigma1", "dsigma2", "drou", "eEnergy")


def gradients(cfg, I1, VV, state, assemble=False):
    """One gradient pass (gqmap_gpu_mixture.m:29-34); with assemble=True also :36-40.  Returns dict (+ 'dalpha')."""
    I1, VV = _f64(I1), _f64(VV)
    M, N, L = cfg.M, cfg.N, cfg.L
    out = {n: np.zeros((M, N, L), order="F") for n in NODE_NAMES}
    out.update({n: np.zeros((M, N, L, 2, 2), order="F") for n in EDGE_NAMES})
    g = QoGrads()
    for n in NODE_NAMES + EDGE_NAMES:
        setattr(g, n, _p(out[n]))
    st = state.c_struct()
    lib().qo_gradients(C.byref(cfg), _p(I1), _p(VV), C.byref(st), C.byref(g))
    if assemble:
        dalpha = np.zeros(L)
        lib().qo_assemble(C.byref(cfg), C.byref(g), _p(dalpha))
        out["dalpha"] = dalpha
    return out


def run(cfg, I1, VV, state, it, its, nsteps):
    """Run up to nsteps iterations of the main loop in place on `state`.
    Returns (n_done, it_next, stopped, Energy[n_done], ptdmu[n_done], ptdsigma[n_done])."""
    I1, VV = _f64(I1), _f64(VV)
    E = np.zeros(nsteps)
    dm = np.zeros(nsteps)
    ds = np.zeros(nsteps)
    itc = C.c_int(it)
    stopped = C.c_int(0)
    st = state.c_struct()
    n = lib().qo_run(C.byref(cfg), _p(I1), _p(VV), C.byref(st), C.byref(itc), C.c_int(its), C.c_int(nsteps),
                     _p(E), _p(dm), _p(ds), C.byref(stopped))
    if n < 0:
        raise RuntimeError("qo_run failed")
    state.T = st.T
    return n, itc.value, bool(stopped.value), E[:n], dm[:n], ds[:n]


def interp2_cubic_refine(V, k):
    """interp2(V,k,'cubic') restated (see qo_interp2_cubic_refine)."""
    V = _f64(V)
    M, N = V.shape
    out = np.zeros(((M - 1) * 2 ** k + 1, (N - 1) * 2 ** k + 1), order="F")
    lib().qo_interp2_cubic_refine(_p(V), C.c_int(M), C.c_int(N), C.c_int(k), _p(out))
    return out


_cont_keep = None


def set_nearest_lookup(I2_cont=None, rfc=6):
    """legacy/gqmap_ctf.m:96: data term = nearest lookup into I2_cont (None switches back to the exact bicubic sample)."""
    global _cont_keep
    if I2_cont is None:
        lib().qo_set_nearest_lookup(None, 0, 0, 0)
        _cont_keep = None
        return
    _cont_keep = _f64(I2_cont)
    lib().qo_set_nearest_lookup(_p(_cont_keep), C.c_int(_cont_keep.shape[0]), C.c_int(_cont_keep.shape[1]), C.c_int(2 ** rfc))


def update_alpha(cfg, state, dalpha, step):
    st = state.c_struct()
    dalpha = np.ascontiguousarray(dalpha, dtype=np.float64)
    lib().qo_update_alpha(C.byref(cfg), C.byref(st), _p(dalpha), C.c_double(step))


def projsplx(y):
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    x = np.zeros_like(y)
    lib().qo_projsplx(_p(y), C.c_int(y.size), _p(x))
    return x


def find_map(alpha, mu_u, sig_u, mu_v, sig_v, nthreads=0):
    """get_map_mex(alf, mu_u, sig_u, mu_v, sig_v) -> map (M,N,2)."""
    mu_u, sig_u, mu_v, sig_v = map(_f64, (mu_u, sig_u, mu_v, sig_v))
    if mu_u.ndim == 2:
        mu_u, sig_u, mu_v, sig_v = (a.reshape(a.shape + (1,), order="F") for a in (mu_u, sig_u, mu_v, sig_v))
    M, N, L = mu_u.shape
    alpha = np.ascontiguousarray(np.asarray(alpha, dtype=np.float64).ravel())
    out = np.zeros((M, N, 2), order="F")
    lib().qo_find_map(_p(alpha), _p(mu_u), _p(sig_u), _p(mu_v), _p(sig_v), C.c_int(M), C.c_int(N), C.c_int(L),
                      _p(out), C.c_int(nthreads))
    return out


def fminbnd_mixture(a, u, o, ax, bx):
    a, u, o = (np.ascontiguousarray(v, dtype=np.float64) for v in (a, u, o))
    fval = C.c_double(0)
    cnt = C.c_int(0)
    x = lib().qo_fminbnd_mixture(_p(a), _p(u), _p(o), C.c_int(a.size), C.c_double(ax), C.c_double(bx),
                                 C.byref(fval), C.byref(cnt))
    return x, fval.value, cnt.value


def profile_logp(cfg, I1, VV, uv):
    I1, VV, uv = _f64(I1), _f64(VV), _f64(uv)
    return lib().qo_profile_logp(C.byref(cfg), _p(I1), _p(VV), _p(uv))


def aepe(cfg, flow_map, tflow, unknown):
    flow_map, tflow = _f64(flow_map), _f64(tflow)
    unk = np.asfortranarray(np.asarray(unknown, dtype=np.uint8))
    return lib().qo_aepe(C.byref(cfg), _p(flow_map), _p(tflow), unk.ctypes.data_as(C.POINTER(C.c_ubyte)))


def flow_to_color(flow, max_flow=-1.0):
    """[img,flo,minu,maxu,minv,maxv,idxUnknown] = flowToColor(flow[,maxFlow])  (legacy/flowToColor.m:1)."""
    flow = _f64(flow)
    M, N, _ = flow.shape
    img = np.zeros((M, N, 3), dtype=np.uint8, order="F")
    flo = np.zeros((M, N, 2), order="F")
    stats = np.zeros(4)
    unk = np.zeros((M, N), dtype=np.uint8, order="F")
    lib().qo_flow_to_color(_p(flow), C.c_int(M), C.c_int(N), C.c_double(max_flow),
                           img.ctypes.data_as(C.POINTER(C.c_ubyte)), _p(flo), _p(stats),
                           unk.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return img, flo, stats[0], stats[1], stats[2], stats[3], unk.astype(bool)
