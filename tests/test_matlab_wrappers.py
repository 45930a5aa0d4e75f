"""The MATLAB wrappers (gqmap-opticalflow_b200/matlab/*.m), EXECUTED.  MATLAB is not available here, so the .m files run under
oracle/mlab/minimat.py (the interpreter that also executes the reference's own .m files, tests/test_minimat.py).
  * CPU: `gqmap_mex` is replaced by a recording stand-in with the stateful semantics of the real gateway, which pins the wrappers'
    own logic -- monitoring cadence of qgmap_chunked.m (it = 1 and every `log_every`, gqmap_gpu_mixture.m:52), history arrays
    prefilled as the reference's (NaN / 0 / NaN), PNG names, handle released exactly once through onCleanup (also on errors),
    the constants gqmap_ctf.m hands to the solver (legacy/gqmap_ctf.m:7,36,48-51), argument checks;
    with the REAL gateway (mex/libmexharness.so) and no GPU the wrappers must surface the library's refusal, not hide it.
  * GPU: wrapper -> real MEX gateway -> libqgmap.so must return what the Python mirror returns on the same options."""
import os

import numpy as np
import pytest

from oracle.mlab.minimat import Interp, MatlabError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MDIR = os.path.join(ROOT, "gqmap-opticalflow_b200", "matlab")


class FakeGateway:
    """gqmap_mex stand-in: same commands and output counts as mex/gqmap_mex.cpp, trivial arithmetic"""

    def __init__(self, M, N, L, stop_at=None, fail_on=None):
        self.M, self.N, self.L, self.stop_at, self.fail_on = M, N, L, stop_at, fail_on
        self.calls, self.it, self.live = [], 1, set()

    def __call__(self, nargout, cmd, *a):
        self.calls.append((cmd, nargout, a))
        if cmd == self.fail_on:
            e = MatlabError("stand-in failure in %s" % cmd)
            e.ident = "qgmap:cuda"
            raise e
        if cmd == "create":
            h = 41.0 + len(self.live)
            self.live.add(h)
            self.it = 1
            return (h,)
        assert a[0] in self.live, "command on a destroyed handle"
        if cmd == "destroy":
            self.live.discard(a[0])
            return ()
        if cmd in ("set_state", "init_state"):
            self.state = a[1]
            return ()
        if cmd == "step":
            n, its = int(a[1]), int(a[2])
            nit = max(0, min(n, its - self.it + 1))
            stopped = 0.0
            if self.stop_at is not None and self.it + nit - 1 >= self.stop_at:
                nit, stopped = self.stop_at - self.it + 1, 1.0
            its_done = np.arange(self.it, self.it + nit, dtype=np.float64).reshape(-1, 1)
            self.it += nit
            return (-its_done, its_done / 10, its_done / 100, float(nit), stopped)[:max(nargout, 1)]
        if cmd == "map":
            return (np.full((self.M, self.N, 2), float(self.it - 1), order="F"),)
        if cmd == "aepe":
            return (100.0 + self.it - 1,)
        if cmd == "logp":
            return (-1000.0 - (self.it - 1),)
        if cmd == "get_state":
            z = np.zeros((self.M, self.N, self.L), order="F")
            return (dict(muu=z + 1, muv=z + 2, sigmau=z + 3, sigmav=z + 4, pn=z, rou=np.zeros((self.M, self.N, self.L, 2, 2), order="F"),
                         alpha=np.full((1, 1, self.L), 1.0 / self.L), w=np.zeros((1, 1, self.L)), T=0.0, it=float(self.it)),)
        raise AssertionError("unknown command %s" % cmd)


def _interp(gw, pngs=None, ftc=None):
    ext = {"gqmap_mex": gw}
    if ftc is not None:
        ext["flowToColor_mex"] = ftc
    return Interp([MDIR], externals=ext, on_imwrite=(lambda img, name: pngs.append(name)) if pngs is not None else None)


def _options(**kw):
    o = dict(K=3.0, L=2.0, its=700.0, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=-3.0, maxu=2.0, minv=-1.5, maxv=4.0)
    o.update(kw)
    return o


def test_wrappers_forward_to_solve():
    for fn, variant, shape in (("gqmap_gpu_mixture", 0.0, (10, 12)), ("gqmap_gpuSuper_mix_entropy", 1.0, (16, 24))):
        seen = []

        def gw(nargout, cmd, *a):
            seen.append((cmd, nargout, a))
            return tuple(float(i) for i in range(6))
        I1 = np.arange(shape[0] * shape[1], dtype=np.float64).reshape(shape, order="F")
        out = _interp(gw).call(fn, _options(), I1, I1 + 1, nargout=6)
        assert out == tuple(float(i) for i in range(6))
        (cmd, nargout, a), = seen
        assert cmd == "solve" and nargout == 6 and a[0] == variant and a[1]["K"] == 3.0 and np.array_equal(a[2], I1) and np.array_equal(a[3], I1 + 1)
    with pytest.raises(MatlabError) as e:                                # gqmap_gpuSuper_mix_entropy.m:11: M = Mo/4 must be whole
        _interp(lambda *a: ()).call("gqmap_gpuSuper_mix_entropy", _options(), np.zeros((10, 12)), np.zeros((10, 12)), nargout=6)
    assert e.value.ident == "qgmap:arg"


@pytest.mark.parametrize("variant", [0, 1])
def test_chunked_loop_cadence_history_and_cleanup(variant):
    Mo, No = (16, 24) if variant else (9, 11)
    M, N = (Mo // 4, No // 4) if variant else (Mo, No)
    gw, pngs = FakeGateway(M, N, 2), []
    tflow, unk = np.zeros((Mo, No, 2), order="F"), np.zeros((Mo, No), bool)
    ftc_in = []
    o = _options(verbose=True, trueFlow=tflow, unknownIdx=unk, dir="out", seed=5.0, temperature=0.2 if variant else 0.0)
    fn = "gqmap_gpuSuper_mix_entropy" if variant else "gqmap_gpu_mixture"
    I = _interp(gw, pngs, ftc=lambda n, flow: (ftc_in.append(np.asarray(flow).shape), np.zeros(np.asarray(flow).shape[:2] + (3,), np.uint8))[1:])
    mu, sigma, alpha, AEPE, Energy, logP = I.call(fn, o, np.zeros((Mo, No)), np.zeros((Mo, No)), nargout=6)
    cmds = [c[0] for c in gw.calls]
    assert cmds[0] == "create" and gw.calls[0][2][0] == float(variant) and cmds[1] == "init_state" and gw.calls[1][2][1] == 5.0
    steps = [int(c[2][1]) for c in gw.calls if c[0] == "step"]
    assert steps == [1, 299, 300, 100] and all(int(c[2][2]) == 700 for c in gw.calls if c[0] == "step")     # it = 1, 300, 600, then the rest
    logged = [1, 300, 600]
    assert np.array_equal(Energy.ravel(), -np.arange(1.0, 701.0))                                            # every iteration's Energy
    assert np.array_equal(np.flatnonzero(~np.isnan(AEPE.ravel())) + 1, logged) and np.array_equal(np.flatnonzero(~np.isnan(logP.ravel())) + 1, logged)
    assert AEPE[299, 0] == 400.0 and logP[599, 0] == -1600.0 and AEPE.shape == (700, 1)
    assert pngs == ["out/1.png", "out/300.png", "out/600.png"]                                               # gqmap_gpu_mixture.m:62
    assert ftc_in == [((Mo - 8, No - 8, 2) if variant else (Mo, No, 2))] * 3                                  # super: repelem(map,4,4)(5:end-4,5:end-4,:)
    assert cmds.count("destroy") == 1 and cmds[-1] == "destroy" and not gw.live                              # onCleanup, after get_state
    assert mu.shape == (M, N, 2, 2) and np.array_equal(mu[..., 0], np.ones((M, N, 2))) and np.array_equal(sigma[..., 1], 4 * np.ones((M, N, 2)))
    assert np.asarray(alpha).shape == (1, 1, 2)


def test_chunked_loop_early_stop_init_state_and_error_cleanup():
    gw = FakeGateway(9, 11, 2, stop_at=450)                               # the reference's `break` (:75) inside a chunk
    init = dict(muu=np.zeros((9, 11, 2)), muv=np.zeros((9, 11, 2)), sigmau=np.ones((9, 11, 2)), sigmav=np.ones((9, 11, 2)),
                pn=np.zeros((9, 11, 2)), rou=np.zeros((9, 11, 2, 2, 2)), w=np.zeros((1, 1, 2)))
    o = _options(verbose=True, init=init, temperature=0.3, log_every=100.0)
    mu, sigma, alpha, AEPE, Energy, logP = _interp(gw).call("gqmap_gpu_mixture", o, np.zeros((9, 11)), np.zeros((9, 11)), nargout=6)
    set_state = [c for c in gw.calls if c[0] == "set_state"]
    assert len(set_state) == 1 and set_state[0][2][1]["T"] == 0.3 and "muu" in set_state[0][2][1] and not [c for c in gw.calls if c[0] == "init_state"]
    assert [int(c[2][1]) for c in gw.calls if c[0] == "step"] == [1, 99, 100, 100, 100, 100]
    assert np.count_nonzero(Energy) == 450 and Energy[450, 0] == 0.0                                         # zeros after the stop, as the reference's
    assert np.array_equal(np.flatnonzero(~np.isnan(logP.ravel())) + 1, [1, 100, 200, 300, 400]) and np.all(np.isnan(AEPE))   # no trueFlow given
    assert not gw.live
    gw = FakeGateway(9, 11, 2, fail_on="map")                             # an error inside the loop still releases the handle
    with pytest.raises(MatlabError) as e:
        _interp(gw).call("gqmap_gpu_mixture", _options(verbose=True), np.zeros((9, 11)), np.zeros((9, 11)), nargout=6)
    assert e.value.ident == "qgmap:cuda" and [c[0] for c in gw.calls][-1] == "destroy" and not gw.live


def test_gqmap_ctf_hands_the_legacy_constants_to_the_solver():
    M, N = 8, 10
    gw = FakeGateway(M, N, 1)
    rng = np.random.default_rng(3)
    GRDT = np.asfortranarray(rng.normal(0, 2, (M, N, 2)))
    draws = iter(np.random.default_rng(8).random((4, M * N)))
    I = Interp([MDIR], externals={"gqmap_mex": gw}, rand=lambda shape: next(draws).reshape(shape, order="F"))
    mu, sigma, rou, AEPE, Energy = I.call("gqmap_ctf", dict(its=40.0, K=3.0, epsn=1e-6, lambdad=1.0, lambdas=5.0), np.zeros((M, N)), np.zeros((M, N)),
                                          GRDT, nargout=5)
    o = gw.calls[0][2][1]
    want = dict(L=1.0, temperature=0.0, step0=0.07, step_tau=np.inf, sigma_step_scale=0.3, sigma_min=0.01, sigma_max=25.0, corr_tor=0.999,
                minu=GRDT[..., 0].min(), maxu=GRDT[..., 0].max(), minv=GRDT[..., 1].min(), maxv=GRDT[..., 1].max())
    assert all(o[k] == v for k, v in want.items()), {k: o[k] for k in want}
    st = [c for c in gw.calls if c[0] == "set_state"][0][2][1]
    d = np.random.default_rng(8).random((4, M * N))
    assert np.array_equal(st["muu"], want["minu"] + d[0].reshape((M, N), order="F") * (want["maxu"] - want["minu"]))       # legacy/gqmap_ctf.m:14
    assert np.array_equal(st["sigmav"], d[3].reshape((M, N), order="F") + 3) and np.all(st["pn"] == 0) and np.asarray(st["rou"]).shape == (M, N, 1, 2, 2)
    assert [(c[0], int(c[2][1]), int(c[2][2])) for c in gw.calls if c[0] == "step"] == [("step", 40, 40)]
    assert mu.shape == (M, N, 2) and sigma.shape == (M, N, 2) and rou.shape == (M, N, 2, 2) and Energy.shape == (40, 1) and not gw.live
    want_aepe = np.mean(np.sqrt((GRDT[1:-1, 1:-1, 0] - 1) ** 2 + (GRDT[1:-1, 1:-1, 1] - 2) ** 2))
    assert abs(AEPE[39, 0] - want_aepe) < 1e-14 and np.all(np.isnan(AEPE[:39]))


# ---- the real gateway -----------------------------------------------------------------------------------------------------------
def _real_gateway(pkg):
    from test_mex_gateway import Mex, MexError, _Struct
    mex = Mex(pkg)

    def conv(v):
        if isinstance(v, _Struct):
            out = {}
            for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou", "w", "alpha", "T", "it"):
                try:
                    out[f] = conv(v[f])
                except KeyError:
                    pass
            return out
        if isinstance(v, np.ndarray):
            return np.asfortranarray(v) if v.size != 1 else (bool(v.ravel()[0]) if v.dtype == np.bool_ else float(v.ravel()[0]))
        return v

    def gw(name):
        def call(nargout, *a):
            try:
                r = mex.call(name, nargout, *a)
            except MexError as e:
                err = MatlabError(str(e))
                err.ident = e.ident
                raise err
            if nargout <= 1:
                return () if (nargout == 0 and (r is None or (isinstance(r, list) and not r))) else (conv(r),)
            return tuple(conv(x) for x in r)
        return call
    return {"gqmap_mex": gw("gqmap_mex"), "flowToColor_mex": gw("flowToColor_mex"), "get_map_mex": gw("get_map_mex")}


def test_wrappers_surface_the_missing_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    I = Interp([MDIR], externals=_real_gateway(pkg))
    for o in (_options(its=5.0), _options(its=5.0, verbose=True)):       # one-call solve and the chunked loop
        with pytest.raises(MatlabError) as e:
            I.call("gqmap_gpu_mixture", o, np.zeros((12, 12)), np.ones((12, 12)), nargout=6)
        assert e.value.ident == "qgmap:cuda" and "no CPU fallback" in str(e.value)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_wrappers_through_the_real_gateway(pkg, variant, tmp_path):
    """matlab/gqmap_gpu_mixture.m (one call and options.verbose = qgmap_chunked.m) -> mex/gqmap_mex.cpp -> libqgmap.so == Python mirror."""
    M, N = (64, 96) if variant else (40, 52)
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N)
    o = _options(its=7.0, seed=9.0, log_every=3.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, temperature=0.2 if variant else 0.0,
                 trueFlow=np.asfortranarray(flow), unknownIdx=np.zeros((M, N), bool))
    po = {k: (int(v) if k in ("K", "L", "its", "seed", "log_every") else v) for k, v in o.items()}
    name = "gqmap_gpuSuper_mix_entropy" if variant else "gqmap_gpu_mixture"
    ref = getattr(pkg, name)(po, I1, I2)
    pngs = []
    I = Interp([MDIR], externals=_real_gateway(pkg), on_imwrite=lambda img, fn: pngs.append(fn))
    got = I.call(name, o, I1, I2, nargout=6)
    for g, w in zip(got, ref):
        assert np.array_equal(np.asarray(g), w, equal_nan=True)
    chunked = I.call(name, dict(o, verbose=True, dir=str(tmp_path)), I1, I2, nargout=6)
    for g, w in zip(chunked, ref):
        assert np.array_equal(np.asarray(g).reshape(w.shape), w, equal_nan=True)
    assert [os.path.basename(p) for p in pngs] == ["1.png", "3.png", "6.png"]


# ---- the reference's own DRIVER SCRIPTS, unmodified, on top of the wrappers -------------------------------------------------------
REF = "/root/reference"


def _driver_interp(pkg, solve):
    """Search path = our wrappers first, then the reference tree: optical_flow.m finds OUR gqmap_gpu_mixture.m and the reference's
    readFlowFile.m; flowToColor_mex is the real Linux MEX gateway; toolbox functions (imread, rgb2gray, imresize) are supplied here."""
    from PIL import Image
    ext = _real_gateway(pkg)
    ext["gqmap_mex"] = solve

    def imread(nargout, path):
        return (np.asfortranarray(np.asarray(Image.open(os.path.join(REF, path)).convert("RGB"))),)

    def imresize(nargout, img, scale):
        assert scale == 1.0                                              # both drivers use scale = 1
        return (img,)
    ext.update(imread=imread, imresize=imresize, rgb2gray=lambda n, rgb: (np.asfortranarray(pkg.rgb2gray(np.asarray(rgb))),))
    I = Interp([MDIR, REF], externals=ext)
    saved = []
    I.on_save = lambda fn, names, ws: saved.append((fn, names, {k: ws[k] for k in names}))
    return I, saved


@pytest.mark.parametrize("script,solver,variant,seqs,K,lambdas,T", [
    ("optical_flow.m", "gqmap_gpu_mixture", 0.0, ["Teddy", "Cones"], 9.0, 5.0, 0.0),
    ("optical_flowSuper.m", "gqmap_gpuSuper_mix_entropy", 1.0, ["Venus", "Hydrangea", "Urban2", "Urban3", "Grove3"], 11.0, 16.0, 0.2)])
def test_reference_driver_scripts_run_unmodified_on_the_wrappers(pkg, script, solver, variant, seqs, K, lambdas, T):
    """The drop-in claim, executed: the reference's experiment drivers run as they are; only the functions behind the names changed."""
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present on this box")
    calls = []

    def solve(nargout, cmd, *a):                                         # stands in for the GPU: record the request, return shaped outputs
        assert cmd == "solve" and nargout == 6
        calls.append(a)
        o, I1 = a[1], np.asarray(a[2])
        its, L = int(o["its"]), int(o["L"])
        b = 4 if a[0] == 1.0 else 1
        shp = (I1.shape[0] // b, I1.shape[1] // b, L, 2)
        return (np.zeros(shp, order="F"), np.ones(shp, order="F"), np.full((1, 1, L), 1.0 / L), np.full((its, 1), np.nan), np.zeros((its, 1)),
                np.full((its, 1), np.nan))
    I, saved = _driver_interp(pkg, solve)
    cwd = os.getcwd()
    try:
        os.chdir(REF)                                                    # the scripts use paths relative to the reference root
        ws = I.run_script(os.path.join(REF, script))
    finally:
        os.chdir(cwd)
    assert len(calls) == len(seqs) == len(saved)
    for (var, o, I1, I2), seq, (fn, names, vals) in zip(calls, seqs, saved):
        assert var == variant and (o["K"], o["L"], o["its"], o["lambdas"], o["lambdad"], o["temperature"], o["epsn"]) == (K, 3.0, 30000.0, lambdas, 1.0, T, 1e-6)
        d = seq if seq != "RubberWhale" else "rubberwhale"
        flow = pkg.readFlowFile(os.path.join(REF, "middlebury", d, "flow10.flo"))
        img, flo, minu, maxu, minv, maxv, unk = pkg.flowToColor_mex(flow)                    # what optical_flow.m:12-13 must have produced
        assert np.array_equal(o["trueFlow"], flo) and np.array_equal(np.asarray(o["unknownIdx"], bool), unk)
        assert (o["minu"], o["maxu"], o["minv"], o["maxv"]) == (minu, maxu, minv, maxv)
        assert np.asarray(I1).shape == flow.shape[:2] and np.asarray(I1).dtype == np.float64 and 0 <= np.asarray(I1).min() and np.asarray(I1).max() <= 255
        assert not np.array_equal(I1, I2)
        assert fn.endswith("/" + seq + ".mat") and set(names) == {"options", "AEPE", "mu", "sigma", "alpha", "Energy", "logP"}     # optical_flow.m:28
        assert vals["options"]["dir"] == fn[:-len("/" + seq + ".mat")]
    assert ws["name"] == seqs[-1]
