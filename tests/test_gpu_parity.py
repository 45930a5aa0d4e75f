"""GPU parity tests: the CUDA path (through the C ABI of libqgmap.so) against the CPU fp64 oracle on identical inputs.

Tolerances (stated per north_star): per-iteration objective within 1e-4 relative; gradients are fp32 evaluations of an
fp64 reference, compared relative to the array's largest magnitude; MAP indices/branches bit-exact (x to TolX=1e-4).
"""
import numpy as np
import pytest

from conftest import make_problem, options_from_cfg, state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["1", "4"], ids=["lanes1", "lanes4"])
def lanes_per_belief(request, monkeypatch):
    """Run every test with both thread mappings of the iteration kernel (one thread / four lanes per belief); without the
    override the library picks by problem size and small test problems would only ever exercise the 4-lane kernel."""
    monkeypatch.setenv("QGMAP_LANES", request.param)
    return request.param

GRAD_RTOL = 3e-4        # of max|array| (fp32 vs fp64, cancellation in the score-function sums)
GRAD_ETOL = 1e-3        # per element: relative to the gradient's own magnitude ...
GRAD_EFLOOR = 2e-4      # ... plus this times the 1/(o*pr) amplification (the same form as _assert_state_close's step bound)


def _rel(a, b):
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-300)


def _interior(a):
    return a[1:-1, 1:-1]


@pytest.mark.parametrize("variant,L,K,T,small", [
    ("full", 1, 3, 0.0, False), ("full", 2, 5, 0.0, True), ("full", 3, 9, 0.3, False), ("full", 2, 4, 0.1, True),
    ("full", 2, 7, 0.0, True), ("full", 1, 11, 0.2, False),
    ("super", 1, 3, 0.2, False), ("super", 3, 5, 0.2, True), ("super", 2, 6, 0.0, True), ("super", 2, 11, 0.1, False),
])
def test_gradients_match_oracle(pkg, O, variant, L, K, T, small):
    sup = variant == "super"
    Mo, No = (72, 100) if not sup else (96, 136)
    cfg, I1, I2, st = make_problem(O, Mo, No, L, K, super=sup, seed=7 + K, T=T, small_sigma=small)
    VV = O.get_vv(I2)
    g = O.gradients(cfg, I1, VV, st, assemble=True)
    with pkg.Solver(options_from_cfg(cfg, T=T), I1, I2, variant=variant) as s:
        s.set_state(state_dict(st), T=T)
        d = s.debug_gradients()
    pairs = [("G_muu", g["dmuu"]), ("G_muv", g["dmuv"]), ("G_sigu", g["dsigmau"]), ("G_sigv", g["dsigmav"]),
             ("dpn", g["dpn"]), ("drou", g["drou"])]
    amp = _fp32_amplification(st)
    akey = {"G_muu": "muu", "G_muv": "muv", "G_sigu": "sigmau", "G_sigv": "sigmav", "dpn": "pn", "drou": "rou"}
    for name, ref in pairs:
        r = _rel(_interior(d[name]), _interior(ref))
        assert r < GRAD_RTOL, (name, r)
        # per element: relative GRAD_ETOL of the gradient itself plus the fp32 floor of the 1/(o*pr)-amplified potentials
        err = np.abs(_interior(d[name]) - _interior(ref))
        tol = GRAD_ETOL * np.abs(_interior(ref)) + GRAD_EFLOOR * _interior(amp[akey[name]])
        ratio = float((err / tol).max())
        print("per-element gradient ratio", variant, L, K, T, name, "%.3f" % ratio)
        assert ratio <= 1.0, (name, ratio, float(err.max()))
    e_ref = g["nEnergy"] + g["eEnergy"].sum(axis=(3, 4))
    da_ref = g["dan"] + g["dae"].sum(axis=(3, 4))
    assert _rel(_interior(d["e_px"]), _interior(e_ref)) < 2e-5
    assert _rel(_interior(d["da_px"]), _interior(da_ref)) < 2e-5
    # borders are never touched (gqmap_gpu_mixture.m:41-46 index M_,N_ only)
    for name in ("G_muu", "dpn", "e_px"):
        assert np.all(d[name][0] == 0) and np.all(d[name][-1] == 0) and np.all(d[name][:, 0] == 0) and np.all(d[name][:, -1] == 0)


def _f32(a):
    return a.astype(np.float32).astype(np.float64)


def _fp32_amplification(before):
    """Per-element amplification of the fp32 evaluation error of the potentials (|f| ~ 1e2, relative 1e-7) by the 1/(o*pr)
    factors of the score-function gradients (gqmap_gpu_mixture.m:93,114)."""
    prn = 1 - before.pn ** 2
    pre = 1 - before.rou ** 2                                             # (M,N,L,e,c)
    amp = {}
    for c, (mu, sg) in enumerate((("muu", before.sigu), ("muv", before.sigv))):
        own = np.minimum(pre[:, :, :, 0, c], pre[:, :, :, 1, c])
        nb = np.minimum(np.roll(pre[:, :, :, 0, c], 1, 0), np.roll(pre[:, :, :, 1, c], 1, 1))
        amp[mu] = 1.0 / (sg * np.minimum(prn, np.minimum(own, nb)))
    amp["sigmau"], amp["sigmav"] = amp["muu"], amp["muv"]
    amp["pn"], amp["rou"] = 1.0 / prn, 1.0 / pre
    return amp


def _assert_state_close(O, cfg, I1, VV, got, ref, before, step, where=""):
    """One ascent step from identical state.  The fp32 kernel reproduces each gradient to ~1e-5..1e-4 RELATIVE (the
    reference's 1/(o*pr) factors, gqmap_gpu_mixture.m:93,114, make gradients of beliefs sitting on the |rho|=1-1e-5 or
    sigma=0.01 clamps huge, so an absolute tolerance is meaningless there): the state error must be a small fraction of
    the step actually taken, step*|G_ref|."""
    g = O.gradients(cfg, I1, VV, before, assemble=True)
    amp = _fp32_amplification(before)
    for name, refa, G in (("muu", ref.muu, g["dmuu"]), ("muv", ref.muv, g["dmuv"]), ("sigmau", ref.sigu, g["dsigmau"]),
                          ("sigmav", ref.sigv, g["dsigmav"]), ("pn", ref.pn, g["dpn"]), ("rou", ref.rou, g["drou"])):
        err = np.abs(got[name] - refa)
        tol = 2e-5 + step * (1e-3 * np.abs(G) + 3e-5 * amp[name])
        assert np.all(err <= tol), (where, name, float((err / tol).max()), float(err.max()))


def _round_state(st):
    """The device keeps the beliefs in fp32: give the oracle the same (fp32-representable) starting point."""
    for f in ("muu", "muv", "sigu", "sigv", "pn", "rou"):
        getattr(st, f)[...] = _f32(getattr(st, f))
    return st


@pytest.mark.parametrize("variant,L,K,T", [("full", 1, 3, 0.0), ("full", 2, 5, 0.2), ("super", 3, 5, 0.2)])
def test_iterations_match_oracle(pkg, O, variant, L, K, T):
    """Free-running window from the reference's own kind of initial state (gqmap_gpu_mixture.m:18-24).

    The ascent is chaotic in its early phase: the fp64 oracle started from a state perturbed by ONE fp32 rounding diverges
    from itself by ~1e-4 in Energy after 10 iterations and ~1e-2 after 20 (tests/test_oracle_kats.py::
    test_dynamics_sensitivity).  So a free-running comparison is only meaningful (a) over the first few iterations at the
    north_star tolerance and (b) afterwards against the oracle's own sensitivity ("shadowing"): the GPU trajectory must
    stay as close to the oracle as an fp32-perturbed copy of the oracle does, within a factor."""
    sup = variant == "super"
    Mo, No = (60, 84) if not sup else (96, 128)
    n = 30
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(Mo, No, seed=77)
    cfg = O.make_config(Mo, No, L, K, super=sup, lambdas=16.0 if sup else 5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
    st = _round_state(O.init_state(cfg, 78, T=T))
    VV = O.get_vv(I2)
    ref = st.copy()
    nd, it, stopped, E, dm, ds = O.run(cfg, I1, VV, ref, 1, 10 ** 6, n)
    pert = st.copy()                      # the oracle's own sensitivity to a relative 2^-24 perturbation
    pert.muu *= (1 + 2.0 ** -24); pert.sigv *= (1 - 2.0 ** -24)
    _, _, _, Ep, _, _ = O.run(cfg, I1, VV, pert, 1, 10 ** 6, n)
    with pkg.Solver(options_from_cfg(cfg, T=T), I1, I2, variant=variant) as s:
        s.set_state(state_dict(st), T=T)
        r = s.step(n, its=10 ** 6)
        got = s.get_state()
    assert r["n_done"] == nd == n and not r["stopped"] and got["it"] == it
    relE = np.abs(r["Energy"] / E - 1)
    assert relE[:2].max() < 1e-5, relE[:2]        # before the chaos sets in
    assert relE.max() < 0.2 and np.all(r["Energy"] < 0)
    # (b) re-synchronised every iteration along the oracle's trajectory: the north_star gate, 1e-4 relative per iteration
    ref = st.copy()
    with pkg.Solver(options_from_cfg(cfg, T=T), I1, I2, variant=variant) as s:
        for k in range(1, n + 1):
            before = _round_state(ref).copy()
            s.set_state(state_dict(before), T=before.T, it=k, alpha=before.alpha)
            _, _, _, E1, dm1, ds1 = O.run(cfg, I1, VV, ref, k, 10 ** 6, 1)
            r1 = s.step(1)
            assert abs(r1["Energy"][0] / E1[0] - 1) < 1e-5, (k, r1["Energy"][0], E1[0])
            assert abs(r1["ptdmu"][0] / dm1[0] - 1) < 1e-3 and abs(r1["ptdsigma"][0] / ds1[0] - 1) < 1e-3
            if k % 10 == 0:
                _assert_state_close(O, cfg, I1, VV, s.get_state(), ref, before, cfg.step0 / (1 + k / cfg.step_tau), where=k)
    # frozen border (gqmap_gpu_mixture.m:41-46): bit-identical to the initial state
    for name, init in (("muu", st.muu), ("sigmav", st.sigv), ("rou", st.rou)):
        for sl in ((0,), (-1,), (slice(None), 0), (slice(None), -1)):
            assert np.array_equal(got[name][sl], init[sl]), name


@pytest.mark.parametrize("variant,L,K", [("full", 2, 3), ("super", 2, 3)])
def test_single_steps_from_identical_state(pkg, O, variant, L, K):
    """Per-iteration parity proper: re-synchronise the GPU with the oracle's state before every step, on a harsh random
    state (small sigmas, correlations up to 0.6) where free-running trajectories are chaotic."""
    sup = variant == "super"
    Mo, No = (48, 64) if not sup else (64, 96)
    cfg, I1, I2, st = make_problem(O, Mo, No, L, K, super=sup, seed=13, small_sigma=True)
    VV = O.get_vv(I2)
    ref = st.copy()
    with pkg.Solver(options_from_cfg(cfg), I1, I2, variant=variant) as s:
        for it in range(1, 9):
            before = _round_state(ref).copy()
            s.set_state(state_dict(before), it=it, alpha=before.alpha)
            _, _, _, E, dm, ds = O.run(cfg, I1, VV, ref, it, 10 ** 6, 1)
            r = s.step(1)
            got = s.get_state()
            assert abs(r["Energy"][0] / E[0] - 1) < 1e-5
            assert abs(r["ptdmu"][0] / dm[0] - 1) < 1e-4
            _assert_state_close(O, cfg, I1, VV, got, ref, before, cfg.step0 / (1 + it / cfg.step_tau), where=it)


def test_stop_rule_and_history(pkg, O):
    """`it > its` break (:75): exactly `its` iterations run, later requests are no-ops."""
    cfg, I1, I2, st = make_problem(O, 40, 52, 2, 3, seed=5)
    with pkg.Solver(options_from_cfg(cfg), I1, I2) as s:
        s.set_state(state_dict(st))
        r1 = s.step(30, its=37)
        r2 = s.step(30, its=37)
        r3 = s.step(5, its=37)
        assert (r1["n_done"], r1["stopped"]) == (30, False)
        assert (r2["n_done"], r2["stopped"]) == (7, True)
        assert (r3["n_done"], r3["stopped"]) == (0, True)
        assert s.get_state()["it"] == 38


def test_alpha_update_matches_oracle(pkg, O):
    """updateAlpha (:78-86) / projsplx (:49) only fire for it>500; start the counter there, re-synchronise every step."""
    for mode in (0, 1):
        cfg, I1, I2, st = make_problem(O, 40, 52, 3, 3, seed=11, small_sigma=True)
        cfg.alpha_mode = mode
        cfg.alpha_scale = 1e-5            # make the update visible in a short window
        VV = O.get_vv(I2)
        ref = st.copy()
        opts = options_from_cfg(cfg, alpha_mode="projsplx" if mode else "softmax", alpha_scale=1e-5)
        with pkg.Solver(opts, I1, I2) as s:
            for it in range(499, 505):
                before = _round_state(ref).copy()
                s.set_state(state_dict(before), it=it, alpha=before.alpha)
                O.run(cfg, I1, VV, ref, it, 10 ** 6, 1)
                s.step(1)
                got = s.get_state()
                if it <= 500:
                    assert np.array_equal(got["alpha"], before.alpha) and np.array_equal(got["w"], before.w)
                else:
                    assert not np.array_equal(got["alpha"], before.alpha)
                assert np.abs(got["alpha"] - ref.alpha).max() < 1e-6 * max(1.0, np.abs(ref.alpha - before.alpha).max() / 1e-3)
                assert abs(got["alpha"].sum() - 1) < 1e-12
                if mode == 0:
                    assert np.abs(got["w"] - ref.w).max() < 1e-6


def test_super_anneal(pkg, O):
    cfg, I1, I2, st = make_problem(O, 48, 64, 1, 3, super=True, seed=2, T=0.2)
    with pkg.Solver(options_from_cfg(cfg, T=0.2, drate=0.75), I1, I2, variant="super") as s:
        s.set_state(state_dict(st), T=0.2, it=498)
        s.step(5)
        assert abs(s.get_state()["T"] - 0.15) < 1e-15        # it=500 annealed once (S:72)


@pytest.mark.parametrize("L", [1, 2, 3, 5])
def test_find_map_matches_oracle(pkg, O, L):
    rng = np.random.default_rng(L)
    M, N = 37, 53
    alpha = rng.random(L); alpha /= alpha.sum()
    mu_u = rng.uniform(-5, 5, (M, N, L)); mu_v = rng.uniform(-2, 2, (M, N, L))
    sg_u = rng.uniform(0.05, 3, (M, N, L)); sg_v = rng.uniform(0.05, 3, (M, N, L))
    mu_v[:5] = mu_v[:5, :, :1]                       # degenerate: all means equal
    ref = O.find_map(alpha, mu_u, sg_u, mu_v, sg_v)
    got = pkg.get_map_mex(alpha, mu_u, sg_u, mu_v, sg_v)
    # identical branch (component mean vs Brent minimiser) everywhere; x within TolX
    is_mean_ref = (ref[:, :, :1] == mu_u).any(axis=2), (ref[:, :, 1:] == mu_v).any(axis=2)
    is_mean_got = (got[:, :, :1] == mu_u).any(axis=2), (got[:, :, 1:] == mu_v).any(axis=2)
    assert np.array_equal(is_mean_ref[0], is_mean_got[0]) and np.array_equal(is_mean_ref[1], is_mean_got[1])
    assert np.abs(got - ref).max() < 1e-4
    assert (np.abs(got - ref) > 1e-12).mean() < 0.01          # almost everywhere bit-close


def test_handle_map_logp_aepe(pkg, O):
    for variant, (Mo, No) in (("full", (44, 60)), ("super", (64, 96))):
        sup = variant == "super"
        cfg, I1, I2, st = make_problem(O, Mo, No, 3, 3, super=sup, seed=9, small_sigma=True)
        VV = O.get_vv(I2)
        rng = np.random.default_rng(1)
        tflow = rng.uniform(-2, 2, (Mo, No, 2))
        unk = rng.random((Mo, No)) < 0.1
        with pkg.Solver(options_from_cfg(cfg), I1, I2, variant=variant) as s:
            s.set_state(state_dict(st))
            m = s.map()
            f32 = lambda a: a.astype(np.float32).astype(np.float64)
            ref = O.find_map(st.alpha, f32(st.muu), f32(st.sigu), f32(st.muv), f32(st.sigv))
            assert np.abs(m - ref).max() < 1e-9
            lp, lp_ref = s.logp(ref), O.profile_logp(cfg, I1, VV, ref)
            assert abs(lp / lp_ref - 1) < 1e-12
            a, a_ref = s.aepe(ref, tflow, unk), O.aepe(cfg, ref, tflow, unk)
            assert abs(a / a_ref - 1) < 1e-12


def test_solve_outputs(pkg, O):
    """[mu,sigma,alpha,AEPE,Energy,logP] shapes / prefill conventions (:16,:183-188) and values vs the oracle loop."""
    Mo, No, L, K, its = 44, 60, 2, 3, 20
    cfg, I1, I2, st = make_problem(O, Mo, No, L, K, seed=4, small_sigma=True)
    st.pn[:] = 0; st.rou[:] = 0
    _round_state(st)
    VV = O.get_vv(I2)
    rng = np.random.default_rng(0)
    tflow = rng.uniform(-1, 1, (Mo, No, 2))
    unk = np.zeros((Mo, No), bool)
    opts = options_from_cfg(cfg, its=its, init=state_dict(st), trueFlow=tflow, unknownIdx=unk, log_every=8)
    mu, sigma, alpha, AEPE, Energy, logP = pkg.gqmap_gpu_mixture(opts, I1, I2)
    assert mu.shape == (Mo, No, L, 2) and sigma.shape == (Mo, No, L, 2) and alpha.shape == (1, 1, L)
    assert AEPE.shape == Energy.shape == logP.shape == (its, 1)
    logged = [0, 7, 15]                                   # it = 1, 8, 16
    assert np.isfinite(AEPE[logged]).all() and np.isnan(np.delete(AEPE, logged)).all()
    assert np.isfinite(logP[logged]).all() and np.isnan(np.delete(logP, logged)).all()
    nl, ms = pkg.last_solve_stats()
    assert nl >= its and ms > 0
    ref = st.copy()
    _, _, _, E, _, _ = O.run(cfg, I1, VV, ref, 1, its, its)
    assert np.abs(Energy[:3, 0] / E[:3] - 1).max() < 1e-5 and np.all(Energy < 0)     # later iterations: chaotic drift
    # one iteration through the one-call solve: every returned belief within the per-element bound of the step it took
    mu1, sigma1, alpha1, _, E1, _ = pkg.gqmap_gpu_mixture(dict(opts, its=1), I1, I2)
    before = _round_state(st.copy())
    ref1 = before.copy()
    O.run(cfg, I1, VV, ref1, 1, 1, 1)
    g = O.gradients(cfg, I1, VV, before, assemble=True)
    amp = _fp32_amplification(before)
    step1 = cfg.step0 / (1 + 1 / cfg.step_tau)
    for name, a, b, G in (("muu", mu1[..., 0], ref1.muu, g["dmuu"]), ("muv", mu1[..., 1], ref1.muv, g["dmuv"]),
                          ("sigmau", sigma1[..., 0], ref1.sigu, g["dsigmau"]), ("sigmav", sigma1[..., 1], ref1.sigv, g["dsigmav"])):
        err = np.abs(a - b)
        tol = 2e-5 + step1 * (1e-3 * np.abs(G) + 3e-5 * amp[name])
        assert np.all(err <= tol), (name, float((err / tol).max()), float(err.max()))
    mu2, sigma2, alpha2, _, E2, _ = pkg.gqmap_gpu_mixture(dict(opts, its=2), I1, I2)
    ref = _round_state(st.copy())
    O.run(cfg, I1, VV, ref, 1, 2, 2)
    assert np.array_equal(E2[:, 0], Energy[:2, 0])                                    # deterministic
    for a, b in ((mu2[..., 0], ref.muu), (mu2[..., 1], ref.muv), (sigma2[..., 0], ref.sigu), (sigma2[..., 1], ref.sigv)):
        assert (np.abs(a - b) > 5e-4).mean() < 0.01 and np.abs(a - b).max() < 0.2
    assert np.allclose(alpha2.ravel(), ref.alpha, atol=1e-15)


def test_programmatic_dependent_launch_is_bit_identical(pkg, O, monkeypatch):
    """QGMAP_PDL=1 (iteration launches carry the programmatic-serialization attribute, the kernel waits on griddepcontrol before it reads
    anything the previous iteration wrote) must not change a single bit, also across the 25-node CUDA graphs."""
    cfg, I1, I2, st = make_problem(O, 60, 84, 2, 5, seed=11)
    out = []
    for pdl in ("0", "1"):
        monkeypatch.setenv("QGMAP_PDL", pdl)
        with pkg.Solver(options_from_cfg(cfg), I1, I2, variant="full") as s:
            s.set_state(state_dict(st))
            r = s.step(60)
            out.append((r["Energy"].copy(), s.get_state()))
    assert np.array_equal(out[0][0], out[1][0])
    for k in ("muu", "muv", "sigmau", "sigmav", "pn", "rou"):
        assert np.array_equal(out[0][1][k], out[1][1][k]), k


def test_single_precision_state_boundary(pkg, O):
    """qgmap_set_state_f32 / qgmap_get_state_f32 (MATLAB `single` arrays at the boundary): the device keeps the beliefs in fp32, so the
    format is lossless -- same trajectory as the double-precision boundary from fp32-representable values, get == get.astype(float32),
    and get_state_f32 -> set_state_f32 restores the state bit for bit."""
    cfg, I1, I2, st = make_problem(O, 44, 60, 2, 5, seed=13)
    sd = state_dict(_round_state(st))
    fields = ("muu", "muv", "sigmau", "sigmav", "pn", "rou")
    with pkg.Solver(options_from_cfg(cfg), I1, I2, variant="full") as s:
        s.set_state(sd)
        r64 = s.step(5)
        a, a32 = s.get_state(), s.get_state(np.float32)
        for k in fields:
            assert a32[k].dtype == np.float32 and a32[k].shape == a[k].shape
            assert np.array_equal(a32[k].astype(np.float64), a[k]), k
        assert a32["it"] == a["it"] and np.array_equal(a32["alpha"], a["alpha"]) and np.array_equal(a32["w"], a["w"])
        s.set_state({k: (np.asarray(v, dtype=np.float32) if k != "w" else v) for k, v in sd.items()})
        r32 = s.step(5)
        b = s.get_state()
        assert np.array_equal(r64["Energy"], r32["Energy"])
        for k in fields:
            assert np.array_equal(a[k], b[k]), k
        s.set_state(a32, T=a32["T"], it=a32["it"], alpha=a32["alpha"])       # checkpoint / resume through the fp32 boundary
        c = s.get_state(np.float32)
        for k in fields:
            assert np.array_equal(a32[k], c[k]), k
        assert c["it"] == a32["it"]
