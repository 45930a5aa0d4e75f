"""The MATLAB-facing boundary, EXECUTED: the three MEX gateways (gqmap_mex.cpp, get_map_mex.cpp, flowToColor_mex.cpp) are linked,
unchanged, against a small stand-in for MATLAB's MEX runtime (mex/stub/mex_runtime.cpp) and called with the mxArrays a MATLAB
caller would pass: `gqmap_mex('solve', variant, options, I1, I2)` is what matlab/gqmap_gpu_mixture.m runs, and
`get_map_mex` / `flowToColor_mex` replace the reference's Windows-only .mexw64 binaries (gqmap_gpu_mixture.m:57,60; optical_flow.m:12).
Results must equal the ctypes path on the same inputs; argument errors must carry the ids of the reference's Coder gateways."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "gqmap-opticalflow_b200", "mex", "libmexharness.so")
DOUBLE, LOGICAL, UINT8, STRUCT = 6, 3, 9, 2


class Mex:
    def __init__(self, pkg):
        assert os.path.exists(HARNESS), "run `make -C gqmap-opticalflow_b200/mex` (build() does)"
        self.pkg = pkg                                    # libqgmap.so is loaded (RTLD_GLOBAL) by the package
        self.lib = C.CDLL(HARNESS)
        L = self.lib
        for n in ("mh_new_double", "mh_new_logical", "mh_new_string", "mh_new_struct", "mh_get_field", "mh_data"):
            getattr(L, n).restype = C.c_void_p
        L.mh_new_double.argtypes = [C.c_int, C.POINTER(C.c_size_t), C.c_void_p]
        L.mh_new_logical.argtypes = [C.c_int, C.POINTER(C.c_size_t), C.c_void_p]
        L.mh_new_string.argtypes = [C.c_char_p]
        L.mh_set_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.mh_get_field.argtypes = [C.c_void_p, C.c_char_p]
        for n in ("mh_class", "mh_ndim"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.mh_dims.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.mh_data.argtypes = [C.c_void_p]
        L.mh_nbytes.argtypes = [C.c_void_p]
        L.mh_nbytes.restype = C.c_size_t
        L.mh_free.argtypes = [C.c_void_p]
        L.mh_call.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_char_p, C.c_char_p, C.c_int]

    def to_mx(self, v):
        L = self.lib
        if isinstance(v, str):
            return L.mh_new_string(v.encode())
        if isinstance(v, dict):
            s = L.mh_new_struct()
            for k, x in v.items():
                L.mh_set_field(s, k.encode(), self.to_mx(x))
            return s
        a = np.asarray(v)
        if a.dtype == bool:
            a = np.asfortranarray(a.astype(np.uint8))
            dims = (C.c_size_t * max(a.ndim, 2))(*(a.shape if a.ndim >= 2 else a.shape + (1,) * (2 - a.ndim)))
            return L.mh_new_logical(max(a.ndim, 2), dims, a.ctypes.data)
        a = np.asfortranarray(a, dtype=np.float64)
        shape = a.shape if a.ndim >= 2 else ((1, 1) if a.ndim == 0 else (a.shape[0], 1))
        dims = (C.c_size_t * len(shape))(*shape)
        return L.mh_new_double(len(shape), dims, a.ctypes.data)

    def from_mx(self, p):
        L = self.lib
        cls = L.mh_class(p)
        if cls == STRUCT:
            return _Struct(self, p)
        nd = L.mh_ndim(p)
        dims = (C.c_size_t * nd)()
        L.mh_dims(p, dims)
        shape = tuple(dims)
        dt = {DOUBLE: np.float64, LOGICAL: np.bool_, UINT8: np.uint8}[cls]
        n = L.mh_nbytes(p)
        buf = (C.c_ubyte * n).from_address(L.mh_data(p)) if n else b""
        return np.frombuffer(bytes(buf), dtype=dt).reshape(shape, order="F").copy()

    def call(self, gateway, nlhs, *args):
        """MATLAB's `[o1,...,o_nlhs] = gateway(args...)`; raises MexError(id, msg) where MATLAB would raise."""
        L = self.lib
        prhs = (C.c_void_p * max(len(args), 1))(*[self.to_mx(a) for a in args])
        plhs = (C.c_void_p * max(nlhs, 1))()
        eid, emsg = C.create_string_buffer(1024), C.create_string_buffer(1024)
        rc = L.mh_call(gateway.encode(), nlhs, plhs, len(args), prhs, eid, emsg, 1024)
        for p in prhs[:len(args)]:
            L.mh_free(p)
        if rc:
            raise MexError(eid.value.decode(), emsg.value.decode())
        out = [self.from_mx(plhs[i]) for i in range(max(nlhs, 1)) if plhs[i]]
        return out[0] if len(out) == 1 else out


class _Struct:
    def __init__(self, mex, p):
        self.mex, self.p = mex, p

    def __getitem__(self, name):
        f = self.mex.lib.mh_get_field(self.p, name.encode())
        if not f:
            raise KeyError(name)
        return self.mex.from_mx(f)


class MexError(RuntimeError):
    def __init__(self, ident, msg):
        super().__init__("%s: %s" % (ident, msg))
        self.ident = ident


@pytest.fixture(scope="module")
def mex(pkg):
    return Mex(pkg)


def test_flowToColor_mex_gateway(pkg, mex):
    """flowToColor_mex is host code: the whole gateway runs without a GPU.  optical_flow.m:12-13 call form (7 outputs)."""
    rng = np.random.default_rng(2)
    flow = rng.normal(0, 3, (21, 34, 2))
    flow[3, 4] = 1.7e9
    img, flo, minu, maxu, minv, maxv, unk = mex.call("flowToColor_mex", 7, flow)
    ref = pkg.flowToColor_mex(flow)
    assert img.dtype == np.uint8 and img.shape == (21, 34, 3) and np.array_equal(img, ref[0])
    assert np.array_equal(flo, ref[1]) and unk.dtype == np.bool_ and np.array_equal(unk, ref[6])
    assert (minu.item(), maxu.item(), minv.item(), maxv.item()) == tuple(ref[2:6])
    only = mex.call("flowToColor_mex", 1, flow, 4.0)                   # gqmap_gpu_mixture.m:60 call form, with maxFlow
    assert np.array_equal(only, pkg.flowToColor_mex(flow, 4.0)[0])
    with pytest.raises(MexError) as e:
        mex.call("flowToColor_mex", 1)
    assert e.value.ident == "EMLRT:runTime:WrongNumberOfInputs"
    with pytest.raises(MexError) as e:
        mex.call("flowToColor_mex", 8, flow)
    assert e.value.ident == "EMLRT:runTime:TooManyOutputArguments"
    with pytest.raises(MexError) as e:
        mex.call("flowToColor_mex", 1, np.zeros((4, 4, 3)))
    assert e.value.ident == "flowToColor:bands"


def test_gateway_argument_errors(pkg, mex):
    with pytest.raises(MexError) as e:                                  # get_map_mex takes exactly five inputs
        mex.call("get_map_mex", 1, np.ones(2), np.zeros((3, 3, 2)))
    assert e.value.ident == "EMLRT:runTime:WrongNumberOfInputs"
    with pytest.raises(MexError) as e:
        mex.call("get_map_mex", 1, np.ones(2), np.zeros((3, 3, 2)), np.ones((3, 4, 2)), np.zeros((3, 3, 2)), np.ones((3, 3, 2)))
    assert e.value.ident == "Coder:MATLAB:catenate_dimensionMismatch"
    with pytest.raises(MexError) as e:                                  # a required field of gqmap_gpu_mixture.m:3-6 is missing
        mex.call("gqmap_mex", 6, "solve", 0.0, dict(K=3.0, L=1.0), np.zeros((8, 8)), np.zeros((8, 8)))
    assert e.value.ident == "qgmap:arg" and "options." in str(e.value)
    with pytest.raises(MexError) as e:
        mex.call("gqmap_mex", 1, "map", 7.0)                            # no such handle
    assert e.value.ident == "qgmap:state"
    with pytest.raises(MexError) as e:
        mex.call("gqmap_mex", 1, "destroy", 7.0)                        # destroy returns nothing
    assert e.value.ident == "EMLRT:runTime:TooManyOutputArguments"
    import torch
    if not torch.cuda.is_available():                                   # no device: the gateway reports the library's refusal
        with pytest.raises(MexError) as e:
            mex.call("get_map_mex", 1, np.ones(1), np.zeros((3, 3, 1)), np.ones((3, 3, 1)), np.zeros((3, 3, 1)), np.ones((3, 3, 1)))
        assert e.value.ident == "qgmap:cuda" and "no CPU fallback" in str(e.value)


def _options(M, N, **kw):
    o = dict(K=3.0, L=2.0, its=7.0, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=-3.0, maxu=2.0, minv=-1.5, maxv=4.0,
             seed=9.0, log_every=3.0)
    o.update(kw)
    return o


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_gqmap_mex_solve_equals_python_path(pkg, mex, variant, tmp_path):
    """[mu,sigma,alpha,AEPE,Energy,logP] = gqmap_mex('solve', variant, options, I1, I2) -- the body of matlab/gqmap_gpu_mixture.m."""
    M, N = (64, 96) if variant else (40, 52)
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N)
    o = _options(M, N, minu=minu, maxu=maxu, minv=minv, maxv=maxv, temperature=0.2 if variant else 0.0,
                 trueFlow=flow, unknownIdx=np.zeros((M, N), bool), alpha_mode="softmax", dir=str(tmp_path))
    mu, sigma, alpha, AEPE, Energy, logP = mex.call("gqmap_mex", 6, "solve", float(variant), o, I1, I2)
    po = {k: (int(v) if k in ("K", "L", "its", "seed", "log_every") else v) for k, v in o.items() if k != "dir"}
    fn = pkg.gqmap_gpuSuper_mix_entropy if variant else pkg.gqmap_gpu_mixture
    ref = fn(po, I1, I2)
    b = 4 if variant else 1
    assert mu.shape == (M // b, N // b, 2, 2) and alpha.shape == (1, 1, 2) and AEPE.shape == (7, 1)      # gqmap_gpu_mixture.m:183-188
    for got, want in zip((mu, sigma, alpha, AEPE, Energy, logP), ref):
        assert np.array_equal(got, want, equal_nan=True)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["1.png", "3.png", "6.png"]                     # options.dir, :59-62
    banded = mex.call("gqmap_mex", 6, "solve", float(variant), dict(o, devices=np.array([0.0, 0.0])), I1, I2)   # options.devices
    assert np.array_equal(banded[0], mu) and np.array_equal(banded[1], sigma)


@pytest.mark.gpu
def test_gqmap_mex_stateful_loop_and_get_map_mex(pkg, mex):
    """The chunked loop of matlab/qgmap_chunked.m over the stateful commands, and the get_map_mex gateway on its output."""
    M, N = 36, 44
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N)
    o = _options(M, N, minu=minu, maxu=maxu, minv=minv, maxv=maxv, L=3.0)
    h = mex.call("gqmap_mex", 1, "create", 0.0, o, I1, I2)
    mex.call("gqmap_mex", 0, "init_state", h, 5.0)
    E, dmu, dsig, nit, stopped = mex.call("gqmap_mex", 5, "step", h, 6.0, 100.0)
    S = mex.call("gqmap_mex", 1, "get_state", h)
    mp = mex.call("gqmap_mex", 1, "map", h)
    lp = mex.call("gqmap_mex", 1, "logp", h, mp)
    ae = mex.call("gqmap_mex", 1, "aepe", h, mp, flow, np.zeros((M, N), bool))
    mex.call("gqmap_mex", 0, "destroy", h)
    po = {k: (int(v) if k in ("K", "L", "its", "seed", "log_every") else v) for k, v in o.items()}
    with pkg.Solver(po, I1, I2) as s:
        s.init_state(5)
        r = s.step(6, its=100)
        st = s.get_state()
        m2 = s.map()
        lp2, ae2 = s.logp(m2), s.aepe(m2, flow, np.zeros((M, N), bool))
    assert int(nit.item()) == 6 and int(stopped.item()) == 0 and np.array_equal(E.ravel()[:6], r["Energy"])
    for k in ("muu", "muv", "sigmau", "sigmav", "pn"):
        assert np.array_equal(S[k], st[k]), k
    assert np.array_equal(mp, m2) and lp.item() == lp2 and ae.item() == ae2
    alf = S["alpha"]
    got = mex.call("get_map_mex", 1, alf, S["muu"], S["sigmau"], S["muv"], S["sigmav"])            # gqmap_gpu_mixture.m:57
    assert got.shape == (M, N, 2) and np.array_equal(got, pkg.get_map_mex(alf.ravel(), S["muu"], S["sigmau"], S["muv"], S["sigmav"]))
