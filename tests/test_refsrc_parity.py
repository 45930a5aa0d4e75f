"""The iteration loop pinned against THE REFERENCE'S OWN SOURCE, executed.

tests/golden/refsrc_*.npz were produced by running gqmap_gpu_mixture.m / gqmap_gpuSuper_mix_entropy.m -- unmodified, read from
/root/reference -- under the mini-MATLAB interpreter oracle/mlab/minimat.py (an independent implementation of MATLAB's language
and built-ins that knows nothing about the algorithm), with the MEX calls routed to the reference's own .mexw64 machine code
(tests/golden/make_refsrc_golden.py).  Each file holds the `rand` draws the program consumed, its outputs, the state it never
returns and probes taken at the solver's own per-iteration fprintf.
  * CPU: the oracle's C restatement, started from the same draws, must reproduce Energy(it), mu, sigma, pn, rou, alpha, AEPE and
    logP of the executed source to fp64 rounding; probes 500 iterations into a run pin the alpha update (it > 500) and the
    temperature anneal (super, every 500) step by step.
  * live (build container only): the interpreter re-runs a case from the reference tree and must reproduce the committed file;
    GaussHermite_2.m, projsplx.m and getVV are executed and compared with the oracle's restatements.
  * GPU: one CUDA step from the same state against what the executed source produced (north_star: objective within 1e-4)."""
import os
import sys

import numpy as np
import pytest

from conftest import options_from_cfg, state_dict

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
SHORT = ("full_L2K3", "full_L1K4", "full_L3K3_T", "super_L2K3_T", "super_L1K3", "full_zero_v", "full_L2K5", "full_L2K9", "full_L1K7")
LONG = ("full_alpha", "super_anneal")
GPU_CASES = SHORT


def load(name):
    d = dict(np.load(os.path.join(GOLD, "refsrc_%s.npz" % name)))
    Mo, No, L, K, T, drate, lambdas, its, minu, maxu, minv, maxv = d["meta"]
    d["cfgargs"] = dict(Mo=int(Mo), No=int(No), L=int(L), K=int(K), super=str(d["solver"]) == "gqmap_gpuSuper_mix_entropy",
                        lambdas=float(lambdas), drate=float(drate), minu=minu, maxu=maxu, minv=minv, maxv=maxv)
    d["T"], d["its"] = float(T), int(its)
    return d


def config(O, d):
    a = dict(d["cfgargs"])
    return O.make_config(a.pop("Mo"), a.pop("No"), a.pop("L"), a.pop("K"), **a)


def initial_state(O, d, cfg):
    """gqmap_gpu_mixture.m:18-24 applied to the draws the executed program consumed (same IEEE operations)."""
    w, ru, rv, su, sv = (np.array(d["draw%d" % i]) for i in range(5))
    M, N, L = cfg.M, cfg.N, cfg.L
    shp = (M, N, L)
    muu = cfg.minu + ru.reshape(shp, order="F") * (cfg.maxu - cfg.minu)
    muv = cfg.minv + rv.reshape(shp, order="F") * (cfg.maxv - cfg.minv)
    sigu = su.reshape(shp, order="F") + (cfg.maxu - cfg.minu)
    sigv = sv.reshape(shp, order="F") + (cfg.maxv - cfg.minv)
    return O.State(muu, muv, sigu, sigv, np.zeros(shp), np.zeros(shp + (2, 2)), np.ravel(w), T=d["T"])


def probe_state(O, d, cfg, it, T):
    shp = (cfg.M, cfg.N, cfg.L)
    g = lambda f: np.array(d["p%d_%s" % (it, f)])              # copies: O.run updates a State in place
    return O.State(g("muu").reshape(shp, order="F"), g("muv").reshape(shp, order="F"), g("sigmau").reshape(shp, order="F"),
                   g("sigmav").reshape(shp, order="F"), g("pn").reshape(shp, order="F"), g("rou").reshape(shp + (2, 2), order="F"),
                   np.ravel(g("w")), alpha=np.ravel(g("alpha")), T=T)


def _close(a, b, tol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a.reshape(-1) - b.reshape(-1)).max() if a.size else 0.0
    assert err <= tol, (what, err)


def test_files_present():
    for name in SHORT + LONG:
        d = load(name)
        assert d["Energy"].size == d["its"] and np.all(np.isfinite(d["Energy"][:int(d["it_end"]) - 1]))
        assert int(d["it_end"]) == d["its"] + 1                         # no early stop: every requested iteration ran


@pytest.mark.parametrize("name", SHORT)
def test_oracle_reproduces_executed_source(O, name):
    d = load(name)
    cfg = config(O, d)
    st = initial_state(O, d, cfg)
    VV = O.get_vv(d["I2"])
    n, it, stopped, E, dm, ds = O.run(cfg, d["I1"], VV, st, 1, d["its"], d["its"])
    assert n == d["its"] and it == int(d["it_end"])
    # Free-running: the two agree to rounding (1e-15) after the first iteration; from then on the ascent amplifies the difference
    # (correlations reach the 1-1e-5 clamp, where gradients carry a factor 1/(1-rho^2) = 5e4), to 1e-8 in the
    # state by iteration 3-4.  The tight per-iteration statement is test_oracle_single_steps_from_probed_states below.
    assert abs(E[0] / d["Energy"][0] - 1) < 1e-13 and np.abs(E / d["Energy"] - 1).max() < 1e-9, np.abs(E / d["Energy"] - 1)
    shp = (cfg.M, cfg.N, cfg.L)
    mu, sg = d["mu"].reshape(shp + (2,), order="F"), d["sigma"].reshape(shp + (2,), order="F")
    _close(st.muu, mu[..., 0], 1e-6, "muu"); _close(st.muv, mu[..., 1], 1e-6, "muv")
    _close(st.sigu, sg[..., 0], 1e-6, "sigmau"); _close(st.sigv, sg[..., 1], 1e-6, "sigmav")
    _close(st.pn, d["pn"], 1e-6, "pn"); _close(st.rou, d["rou"], 1e-6, "rou")
    _close(st.alpha, d["alpha"], 1e-15, "alpha"); _close(st.w, d["w"], 0.0, "w")
    for k in d["probes"]:                                                # mean|dmu|, mean|dsigma| the solver printed
        assert abs(dm[k - 1] / d["p%d_ptdmu" % k] - 1) < 1e-7 and abs(ds[k - 1] / d["p%d_ptdsigma" % k] - 1) < 1e-7
    # monitoring at it == 1 (:52-67): MAP (the reference's binary), AEPE, profile_logP on the state after the first update
    p1 = probe_state(O, d, cfg, 1, d["T"])
    if cfg.L == 1:
        mp = np.concatenate([p1.muu, p1.muv], axis=2)
    else:
        mp = O.find_map(p1.alpha, p1.muu, p1.sigu, p1.muv, p1.sigv)
    assert abs(O.aepe(cfg, mp, d["tflow"], d["unknown"]) / d["AEPE"][0] - 1) < 1e-12
    assert abs(O.profile_logp(cfg, d["I1"], VV, mp) / d["logP"][0] - 1) < 1e-12
    assert np.all(np.isnan(d["AEPE"][1:])) and np.all(np.isnan(d["logP"][1:]))


@pytest.mark.parametrize("name", SHORT + LONG)
def test_oracle_single_steps_from_probed_states(O, name):
    """Iteration k+1 from the state the executed source held after iteration k.  The long runs reach the alpha update
    (gqmap_gpu_mixture.m:50,78-86: only for it > 500) and the temperature anneal (gqmap_gpuSuper_mix_entropy.m:72: it = 500)."""
    d = load(name)
    cfg = config(O, d)
    VV = O.get_vv(d["I2"])
    probes = [int(k) for k in d["probes"]]
    done = 0
    for k in probes:
        if k + 1 not in probes:
            continue
        T_next = float(d["p%d_T" % (k + 1)])
        st = probe_state(O, d, cfg, k, T_next)
        before_alpha = st.alpha.copy()
        n, it, stopped, E, dm, ds = O.run(cfg, d["I1"], VV, st, k + 1, 10 ** 6, 1)
        ref = probe_state(O, d, cfg, k + 1, T_next)
        assert abs(E[0] / float(d["p%d_Energy" % (k + 1)]) - 1) < 1e-12, (k, E[0])
        assert abs(dm[0] / float(d["p%d_ptdmu" % (k + 1)]) - 1) < 1e-10
        # fp64 rounding (1e-16 relative on potentials ~1e2) x the 1/(1-rho^2) <= 5e4 factor of the gradients x step 0.1 ~ 5e-11
        for f in ("muu", "muv", "sigu", "sigv", "pn", "rou"):
            _close(getattr(st, f), getattr(ref, f), 1e-9, (k, f))
        _close(st.alpha, ref.alpha, 1e-15, (k, "alpha")); _close(st.w, ref.w, 1e-15, (k, "w"))
        if k + 1 > 500 and cfg.L > 1:
            assert not np.array_equal(st.alpha, before_alpha)          # the update really fired
        elif k + 1 <= 500:
            assert np.array_equal(ref.alpha, before_alpha)
        if k + 2 in probes:                                             # T in effect for the following iteration (anneal at it % 500 == 0)
            assert st.T == float(d["p%d_T" % (k + 2)]), (k, st.T)
        done += 1
    assert done >= 1
    if name == "super_anneal":
        assert float(d["p500_T"]) == 0.2 and abs(float(d["p501_T"]) - 0.15) < 1e-16
    if name == "full_alpha":
        assert not np.array_equal(d["p501_alpha"], d["p500_alpha"]) and np.array_equal(d["p500_alpha"], d["p499_alpha"])


def test_reference_source_live(O):
    sys.path.insert(0, GOLD)
    import make_refsrc_golden as G
    if not os.path.isdir(G.REF):
        pytest.skip("reference tree not present on this box (the committed vectors cover it)")
    from oracle.mlab.minimat import Interp
    # the committed file is what the reference's source produces today
    d = load("full_L2K3")
    out = G.run_case("full_L2K3", its=2, probes=(1, 2))
    assert np.array_equal(out["Energy"], d["Energy"][:2]) and np.array_equal(out["p2_muu"], d["p2_muu"])
    assert out["AEPE"][0] == d["AEPE"][0] and out["logP"][0] == d["logP"][0]
    interp = Interp([G.REF])
    rng = np.random.default_rng(3)
    for K in range(2, 14):                                              # GaussHermite_2.m:21-32
        x, w = interp.call("GaussHermite_2", float(K), nargout=2)
        xo, wo = O.gauss_hermite(K)
        assert np.abs(np.ravel(x) - xo).max() < 1e-13 and np.abs(np.ravel(w) - wo).max() < 1e-14
    for L in (2, 3, 5, 8):                                              # projsplx.m:15-32
        for _ in range(5):
            y = rng.normal(0.3, 0.6, L)
            _close(interp.call("projsplx", y.reshape(1, 1, L)), O.projsplx(y), 1e-15, "projsplx")
    for shape in ((5, 6), (9, 4), (12, 17)):                            # getVV, gqmap_gpu_mixture.m:191-208 (bit-exact)
        V = np.asfortranarray(rng.random(shape) * 255)
        for f in ("gqmap_gpu_mixture", "gqmap_gpuSuper_mix_entropy"):
            assert np.array_equal(interp.call(f, V, local="getVV"), O.get_vv(V))


# ---- BASELINE configs[0] on the reference's own data: RubberWhale crop, L=1, K=3, executed source --------------------------
def _rubberwhale(O):
    sys.path.insert(0, GOLD)
    import make_refsrc_golden as G
    d = np.load(os.path.join(GOLD, "refsrc_rubberwhale_L1K3.npz"))
    I1, I2, opts = G.rubberwhale_inputs()
    draws = G.rubberwhale_draws()
    assert np.array_equal(np.array([x.sum() for x in draws]), d["draws_checksum"])        # same NumPy stream as when the file was made
    cfg = O.make_config(96, 128, 1, 3, lambdas=5.0, epsn=opts["epsn"], minu=opts["minu"], maxu=opts["maxu"], minv=opts["minv"], maxv=opts["maxv"])
    w, ru, rv, su, sv = draws
    shp = (96, 128, 1)
    st = O.State(cfg.minu + ru.reshape(shp) * (cfg.maxu - cfg.minu), cfg.minv + rv.reshape(shp) * (cfg.maxv - cfg.minv),
                 su.reshape(shp) + (cfg.maxu - cfg.minu), sv.reshape(shp) + (cfg.maxv - cfg.minv), np.zeros(shp), np.zeros(shp + (2, 2)), np.ravel(w))
    return G, d, cfg, I1, I2, opts, st


def _rw_check(d, k, st, tol, stride):
    for f, name in (("muu", "muu"), ("muv", "muv"), ("sigu", "sigmau"), ("sigv", "sigmav"), ("pn", "pn"), ("rou", "rou")):
        a = getattr(st, f)
        ref = d["p%d_%s" % (k, name)]
        _close(a[::stride, ::stride].reshape(ref.shape), ref, tol, (k, f))
    sums = np.array([getattr(st, f).sum() for f in ("muu", "muv", "sigu", "sigv", "pn", "rou")] +
                    [(getattr(st, f) ** 2).sum() for f in ("muu", "muv", "sigu", "sigv", "pn", "rou")])
    assert np.abs(sums - d["p%d_sums" % k]).max() <= tol * 96 * 128 * 40, (k, np.abs(sums - d["p%d_sums" % k]).max())


def test_oracle_reproduces_executed_source_on_rubberwhale(O):
    G, d, cfg, I1, I2, opts, st = _rubberwhale(O)
    VV = O.get_vv(I2)
    n, it, stopped, E, dm, ds = O.run(cfg, I1, VV, st.copy(), 1, G.RW_ITS, G.RW_ITS)
    assert abs(E[0] / d["Energy"][0] - 1) < 1e-13 and np.abs(E / d["Energy"] - 1).max() < 1e-9, np.abs(E / d["Energy"] - 1)
    one = st.copy()
    O.run(cfg, I1, VV, one, 1, 10 ** 6, 1)
    _rw_check(d, 1, one, 1e-10, G.RW_STRIDE)                               # the whole state after the first iteration
    assert abs(dm[0] / float(d["p1_ptdmu"]) - 1) < 1e-12 and abs(ds[0] / float(d["p1_ptdsigma"]) - 1) < 1e-12
    mp = np.concatenate([one.muu, one.muv], axis=2)                          # L == 1: map = cat(3, mu_u, mu_v) (:54-55)
    assert abs(O.aepe(cfg, mp, opts["trueFlow"], opts["unknownIdx"]) / d["AEPE"][0] - 1) < 1e-12
    assert abs(O.profile_logp(cfg, I1, VV, mp) / d["logP"][0] - 1) < 1e-12


@pytest.mark.gpu
def test_cuda_step_against_executed_source_on_rubberwhale(pkg, O):
    from test_gpu_parity import _round_state
    G, d, cfg, I1, I2, opts, st = _rubberwhale(O)
    before = _round_state(st)
    with pkg.Solver(options_from_cfg(cfg), I1, I2) as s:
        s.set_state(state_dict(before), it=1, alpha=before.alpha)
        r = s.step(1)
        got = s.get_state()
        mp = s.map()
        lp, ae = s.logp(mp), s.aepe(mp, opts["trueFlow"], opts["unknownIdx"])
    assert abs(r["Energy"][0] / d["Energy"][0] - 1) < 1e-5, (r["Energy"][0], d["Energy"][0])       # north_star: 1e-4
    assert abs(r["ptdmu"][0] / float(d["p1_ptdmu"]) - 1) < 1e-4 and abs(r["ptdsigma"][0] / float(d["p1_ptdsigma"]) - 1) < 1e-4
    assert abs(ae / d["AEPE"][0] - 1) < 1e-5 and abs(lp / d["logP"][0] - 1) < 1e-5
    sgot = dict(muu=got["muu"], muv=got["muv"], sigmau=got["sigmau"], sigmav=got["sigmav"])
    for f in sgot:                                                          # first step from sigma ~ 7: gradients O(1..30), fp32 evaluation
        ref = d["p1_%s" % f]
        err = np.abs(sgot[f][::G.RW_STRIDE, ::G.RW_STRIDE].reshape(ref.shape) - ref)
        assert err.max() < 2e-4 and np.median(err) < 2e-6, (f, float(err.max()), float(np.median(err)))


# ---- the same at FULL size: the whole 388 x 584 RubberWhale pair, one iteration of the executed source ---------------------------------
def _rubberwhale_full(O):
    d = np.load(os.path.join(GOLD, "refsrc_rubberwhale_full_L1K3.npz"))
    I1, I2 = np.asfortranarray(d["I1"].astype(np.float64)), np.asfortranarray(d["I2"].astype(np.float64))
    Mo, No = I1.shape
    assert (Mo, No) == (388, 584)
    rng = np.random.default_rng(int(d["seed"]))
    draws = [rng.random(n).reshape(shp, order="F") for n, shp in ((1, (1, 1)),) + ((Mo * No, (Mo, No)),) * 4]
    assert np.array_equal(np.array([x.sum() for x in draws]), d["draws_checksum"])
    minu, maxu, minv, maxv = (float(x) for x in d["range"])
    cfg = O.make_config(Mo, No, 1, 3, lambdas=5.0, epsn=0.001 ** 2, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
    w, ru, rv, su, sv = draws
    shp = (Mo, No, 1)
    st = O.State(minu + ru.reshape(shp) * (maxu - minu), minv + rv.reshape(shp) * (maxv - minv), su.reshape(shp) + (maxu - minu),
                 sv.reshape(shp) + (maxv - minv), np.zeros(shp), np.zeros(shp + (2, 2)), np.ravel(w))
    return d, cfg, I1, I2, st


def test_oracle_reproduces_executed_source_on_full_rubberwhale(O):
    d, cfg, I1, I2, st = _rubberwhale_full(O)
    VV = O.get_vv(I2)
    n, it, stopped, E, dm, ds = O.run(cfg, I1, VV, st, 1, 10 ** 6, 1)
    assert abs(E[0] / d["Energy"][0] - 1) < 1e-13, (E[0], d["Energy"][0])
    assert abs(dm[0] / float(d["p1_ptdmu"]) - 1) < 1e-12 and abs(ds[0] / float(d["p1_ptdsigma"]) - 1) < 1e-12
    for f, name in (("muu", "muu"), ("muv", "muv"), ("sigu", "sigmau"), ("sigv", "sigmav"), ("pn", "pn"), ("rou", "rou")):
        ref = d["p1_" + name]
        _close(getattr(st, f)[::8, ::8].reshape(ref.shape), ref, 1e-10, f)
    sums = np.array([getattr(st, f).sum() for f in ("muu", "muv", "sigu", "sigv", "pn", "rou")] +
                    [(getattr(st, f) ** 2).sum() for f in ("muu", "muv", "sigu", "sigv", "pn", "rou")])
    assert np.abs(sums / d["p1_sums"] - 1)[np.abs(d["p1_sums"]) > 0].max() < 1e-12
    mp = np.concatenate([st.muu, st.muv], axis=2)
    assert abs(O.profile_logp(cfg, I1, VV, mp) / d["logP"][0] - 1) < 1e-12


@pytest.mark.gpu
def test_cuda_step_against_executed_source_on_full_rubberwhale(pkg, O):
    from test_gpu_parity import _round_state
    d, cfg, I1, I2, st = _rubberwhale_full(O)
    before = _round_state(st)
    with pkg.Solver(options_from_cfg(cfg), I1, I2) as s:
        s.set_state(state_dict(before), it=1, alpha=before.alpha)
        r = s.step(1)
        got = s.get_state()
        lp = s.logp(s.map())
    assert abs(r["Energy"][0] / d["Energy"][0] - 1) < 1e-5, (r["Energy"][0], d["Energy"][0])                 # north_star: 1e-4
    assert abs(r["ptdmu"][0] / float(d["p1_ptdmu"]) - 1) < 1e-4 and abs(r["ptdsigma"][0] / float(d["p1_ptdsigma"]) - 1) < 1e-4
    assert abs(lp / d["logP"][0] - 1) < 1e-5
    for f in ("muu", "muv", "sigmau", "sigmav"):
        ref = d["p1_" + f]
        err = np.abs(got[f][::8, ::8].reshape(ref.shape) - ref)
        assert err.max() < 5e-4 and np.median(err) < 2e-6, (f, float(err.max()), float(np.median(err)))


# ---- executed-source cases on the reference's OWN data (tests/golden/make_refsrc_golden.py::REAL_CASES) -----------------------------
# name -> (super-pixel variant, L, K, lambdas, T, drate, frame shape, probe stride, iterations); frames = MATLAB rgb2gray of the shipped
# PNGs (integer grey levels: what the CUDA path gathers from its fp16 one-sector layout while the beliefs are wide), ground truth
# and clamp range from the shipped .flo files through the reference's flowToColor_mex binary
REAL = {
    # BASELINE configs[2] at FULL size: the whole 480 x 640 Urban2 pair, super-pixel variant, constants of optical_flowSuper.m:16-23
    "urban2_super_full_L3K5": (True, 3, 5, 16.0, 0.2, 0.75, (480, 640), 4, 1),
    # the metric's own instantiation (BASELINE configs[3]/[4]: full resolution, L=3, K=5): a window of Grove2, optical_flow.m:16-23
    "grove2_window_L3K5": (False, 3, 5, 5.0, 0.0, 0.5, (128, 160), 4, 2),
    # BASELINE configs[1] (the eight ground-truth sequences, L=2, driver default K=9): a window of Dimetrodon
    "dimetrodon_window_L2K9": (False, 2, 9, 5.0, 0.0, 0.5, (96, 128), 4, 1),
    # optical_flow.m AS SHIPPED (Teddy, K=9, L=3): u range [-52.75, 0], v range [0, 0]
    "teddy_window_L3K9": (False, 3, 9, 5.0, 0.0, 0.5, (96, 128), 4, 1),
    # optical_flowSuper.m AS SHIPPED (Venus, K=11, L=3, super-pixel variant): v range [0, 0]
    "venus_super_window_L3K11": (True, 3, 11, 16.0, 0.2, 0.75, (128, 160), 2, 1),
}
REAL_FILE = {name: os.path.join(GOLD, "refsrc_%s.npz" % name) for name in REAL}
_REAL_FIELDS = (("muu", "muu"), ("muv", "muv"), ("sigu", "sigmau"), ("sigv", "sigmav"), ("pn", "pn"), ("rou", "rou"))


def _real_case(O, name):
    sup, L, K, lambdas, T, drate, shape, stride, its = REAL[name]
    d = np.load(REAL_FILE[name])
    I1, I2 = np.asfortranarray(d["I1"].astype(np.float64)), np.asfortranarray(d["I2"].astype(np.float64))
    Mo, No = I1.shape
    M, N = (Mo // 4, No // 4) if sup else (Mo, No)
    rng = np.random.default_rng(int(d["seed"]))
    draws = [rng.random(n).reshape(shp, order="F") for n, shp in ((L, (1, 1, L)),) + ((M * N * L, (M, N, L)),) * 4]
    assert np.array_equal(np.array([x.sum() for x in draws]), d["draws_checksum"])
    minu, maxu, minv, maxv = (float(x) for x in d["range"])
    cfg = O.make_config(Mo, No, L, K, super=sup, lambdas=lambdas, epsn=0.001 ** 2, minu=minu, maxu=maxu, minv=minv, maxv=maxv, drate=drate)
    w, ru, rv, su, sv = draws
    shp = (M, N, L)
    st = O.State(minu + ru * (maxu - minu), minv + rv * (maxv - minv), su + (maxu - minu), sv + (maxv - minv), np.zeros(shp),
                 np.zeros(shp + (2, 2)), np.ravel(w), T=T)                      # gqmap_gpu_mixture.m:18-24
    return d, cfg, I1, I2, st


def _real_truth(O, name, d, shape):
    """Ground truth of the case: stored for windows; for the whole Urban2 pair read from the reference tree where it exists."""
    if "tflow" in d.files:
        return np.asfortranarray(d["tflow"].astype(np.float64)), d["unknown"]
    gt = "/root/reference/middlebury/Urban2/flow10.flo"
    if not os.path.exists(gt):
        return None, None
    with open(gt, "rb") as f:
        f.read(12)
        flo = np.fromfile(f, np.float32).reshape(480, 640, 2).astype(np.float64)
    r = O.flow_to_color(np.asfortranarray(flo))
    assert tuple(float(x) for x in r[2:6]) == tuple(d["range"])
    return np.asfortranarray(r[1][:shape[0], :shape[1]]), np.asarray(r[6])[:shape[0], :shape[1]]


@pytest.mark.parametrize("name", sorted(REAL))
def test_oracle_reproduces_executed_source_on_reference_data(O, name):
    sup, L, K, lambdas, T, drate, shape, stride, its = REAL[name]
    d, cfg, I1, I2, st = _real_case(O, name)
    VV = O.get_vv(I2)
    assert np.abs(st.alpha - d["alpha"]).max() < 1e-15                        # it <= 500: alpha is still softmax(w) of the draw
    for it in range(1, its + 1):
        tolE, tolS = (1e-13, 1e-10) if it == 1 else (1e-11, 1e-8)              # free-running: iteration 2 inherits the rounding of iteration 1
        n, _, stopped, E, dm, ds = O.run(cfg, I1, VV, st, it, 10 ** 6, 1)
        assert n == 1 and abs(E[0] / d["Energy"][it - 1] - 1) < tolE, (it, E[0], d["Energy"][it - 1])
        assert abs(dm[0] / float(d["p%d_ptdmu" % it]) - 1) < 1e3 * tolE and abs(ds[0] / float(d["p%d_ptdsigma" % it]) - 1) < 1e3 * tolE
        assert st.T == float(d["p%d_T" % it]) == T
        for f, fname in _REAL_FIELDS:
            ref = d["p%d_%s" % (it, fname)]
            _close(getattr(st, f)[::stride, ::stride].reshape(ref.shape), ref, tolS, (it, f))
        sums = np.array([getattr(st, f).sum() for f, _ in _REAL_FIELDS] + [(getattr(st, f) ** 2).sum() for f, _ in _REAL_FIELDS])
        ok = np.abs(d["p%d_sums" % it]) > 0
        assert np.abs(sums[ok] / d["p%d_sums" % it][ok] - 1).max() < 1e2 * tolS
        if it == 1:
            # monitoring at it == 1 (:52-67): get_map_mex (the reference's binary when the file was made; super: repelem(map,4,4) and
            # the 5:end-4 crop), AEPE against the shipped ground truth, profile_logP.  The state agrees to ~1e-13 here; fminbnd's
            # TolX = 1e-4 path and sums of both signs leave 1e-12 .. 1e-11 in the scalars.
            mp = O.find_map(st.alpha, st.muu, st.sigu, st.muv, st.sigv)
            assert abs(O.profile_logp(cfg, I1, VV, mp) / d["logP"][0] - 1) < 1e-10
            tflow, unk = _real_truth(O, name, d, I1.shape)
            if tflow is not None:
                assert abs(O.aepe(cfg, mp, tflow, unk) / d["AEPE"][0] - 1) < 1e-10
    assert np.all(np.isnan(d["AEPE"][1:])) and np.all(np.isnan(d["logP"][1:]))
    assert I1.shape == shape and np.array_equal(I1, np.round(I1)) and 0 <= I1.min() and I1.max() <= 255     # the size claimed; grey levels


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(REAL))
def test_cuda_steps_against_executed_source_on_reference_data(pkg, O, name):
    from test_gpu_parity import _round_state
    from test_gpu_full_size import _assert_step_close
    sup, L, K, lambdas, T, drate, shape, stride, its = REAL[name]
    d, cfg, I1, I2, st = _real_case(O, name)
    VV = O.get_vv(I2)
    tflow, unk = _real_truth(O, name, d, I1.shape)
    before = _round_state(st)
    ref = before.copy()
    O.run(cfg, I1, VV, ref, 1, 10 ** 6, 1)                                     # the oracle from the fp32-rounded start, whole state
    mp_ref = O.find_map(ref.alpha, ref.muu, ref.sigu, ref.muv, ref.sigv)
    with pkg.Solver(options_from_cfg(cfg, T=T), I1, I2, variant="super" if sup else "full") as s:
        s.set_state(state_dict(before), T=T, it=1, alpha=before.alpha)
        r = s.step(1)
        got = s.get_state()
        own = s.map()
        lp_own, lp_on_ref = s.logp(own), s.logp(mp_ref)
        ae_own = s.aepe(own, tflow, unk) if tflow is not None else None
        r2 = s.step(1) if its > 1 else None                                    # free-running second iteration
    print("%s: Energy rel %.2e, ptdmu rel %.2e, ptdsigma rel %.2e, logP(own map) rel %.2e, AEPE diff %s" % (
        name, r["Energy"][0] / d["Energy"][0] - 1, r["ptdmu"][0] / float(d["p1_ptdmu"]) - 1, r["ptdsigma"][0] / float(d["p1_ptdsigma"]) - 1,
        lp_own / d["logP"][0] - 1, None if ae_own is None else "%.2e" % (ae_own - d["AEPE"][0])))
    assert abs(r["Energy"][0] / d["Energy"][0] - 1) < 1e-5, (r["Energy"][0], d["Energy"][0])                 # north_star: 1e-4
    # mean|G| (:69-70): 1e-4 relative plus the fp32 floor of tests/test_gpu_full_size.py (x16 pixels per super-pixel block)
    floor_u = 3e-5 * (16 if sup else 1) * float(np.mean(1.0 / before.sigu[1:-1, 1:-1]))
    assert abs(r["ptdmu"][0] - float(d["p1_ptdmu"])) < 1e-4 * float(d["p1_ptdmu"]) + floor_u
    assert abs(r["ptdsigma"][0] - float(d["p1_ptdsigma"])) < 1e-4 * float(d["p1_ptdsigma"]) + floor_u
    assert abs(lp_on_ref / O.profile_logp(cfg, I1, VV, mp_ref) - 1) < 1e-12   # the logP kernel on the oracle's map
    assert abs(lp_own / d["logP"][0] - 1) < 1e-4                               # the CUDA path's own MAP of its own fp32 state
    if ae_own is not None:
        assert abs(ae_own - d["AEPE"][0]) < 1e-3                               # north_star: EPE within 1e-3 px
    _assert_step_close(got, ref, before, cfg.step0 / (1 + 1 / cfg.step_tau), where=name)
    for f in ("muu", "muv", "sigmau", "sigmav"):                                # and against the executed source itself, on its probes
        pr = d["p1_" + f]
        err = np.abs(got[f][::stride, ::stride].reshape(pr.shape) - pr)
        assert err.max() < 5e-4 and np.median(err) < 2e-6, (f, float(err.max()), float(np.median(err)))
    if r2 is not None:
        assert abs(r2["Energy"][0] / d["Energy"][1] - 1) < 1e-4, (r2["Energy"][0], d["Energy"][1])
        assert abs(r2["ptdmu"][0] / float(d["p2_ptdmu"]) - 1) < 1e-3


# ---- single steps from LATER states of the reference's own trajectory on its own data (at iteration 1 every correlation is still zero) ----
# tests/golden/make_refsrc_golden.py::REAL_LONG: 48 x 64 window of Grove2, L=3, K=5, 16 iterations of gqmap_gpu_mixture.m executed; whole state
# kept after iterations 15 and 16
# name -> (super-pixel variant, L, K, lambdas, T, drate, iteration whose state the step starts from, frame shape)
REAL_LATE = {
    "grove2_window_L3K5_it16": (False, 3, 5, 5.0, 0.0, 0.5, 15, (48, 64)),
    # the super-pixel variant (64 x 80 window of Venus = 16 x 20 beliefs, v range [0, 0], T = 0.2) and the driver's K=9 (32 x 48 window of Dimetrodon, L=2)
    "venus_super_window_L3K5_it12": (True, 3, 5, 16.0, 0.2, 0.75, 11, (64, 80)),
    "dimetrodon_window_L2K9_it8": (False, 2, 9, 5.0, 0.0, 0.5, 7, (32, 48)),
}
LATE_FILE = {name: os.path.join(GOLD, "refsrc_%s.npz" % name) for name in REAL_LATE}


def _long_case(O, name="grove2_window_L3K5_it16"):
    sup, L, K, lambdas, T, drate, k, shape = REAL_LATE[name]
    d = np.load(LATE_FILE[name])
    I1, I2 = np.asfortranarray(d["I1"].astype(np.float64)), np.asfortranarray(d["I2"].astype(np.float64))
    Mo, No = I1.shape
    minu, maxu, minv, maxv = (float(x) for x in d["range"])
    cfg = O.make_config(Mo, No, L, K, super=sup, lambdas=lambdas, epsn=0.001 ** 2, minu=minu, maxu=maxu, minv=minv, maxv=maxv, drate=drate)

    def state(it):
        g = lambda f: np.array(d["f%d_%s" % (it, f)], order="F")
        return O.State(g("muu"), g("muv"), g("sigmau"), g("sigmav"), g("pn"), g("rou"), np.ravel(g("w")), alpha=np.ravel(g("alpha")), T=T)
    return d, cfg, I1, I2, state


@pytest.mark.parametrize("name", sorted(REAL_LATE))
def test_oracle_single_step_from_late_state_on_reference_data(O, name):
    sup, L, K, lambdas, T, drate, k, shape = REAL_LATE[name]
    d, cfg, I1, I2, state = _long_case(O, name)
    VV = O.get_vv(I2)
    st, ref = state(k), state(k + 1)
    assert np.abs(st.pn).max() > 0.01 and np.abs(st.rou).max() > 0.5          # the correlation terms are exercised (zero at iteration 1)
    n, _, stopped, E, dm, ds = O.run(cfg, I1, VV, st, k + 1, 10 ** 6, 1)
    assert n == 1 and abs(E[0] / d["Energy"][k] - 1) < 1e-12, (E[0], d["Energy"][k])
    assert abs(dm[0] / float(d["p%d_ptdmu" % (k + 1)]) - 1) < 1e-10 and abs(ds[0] / float(d["p%d_ptdsigma" % (k + 1)]) - 1) < 1e-10
    assert st.T == float(d["p%d_T" % (k + 1)]) == T
    for f in ("muu", "muv", "sigu", "sigv", "pn", "rou"):                      # fp64 rounding x the 1/(1-rho^2) factor of the gradients x step
        _close(getattr(st, f), getattr(ref, f), 1e-9, f)
    _close(st.alpha, ref.alpha, 1e-15, "alpha")
    assert I1.shape == shape and d["Energy"].size == k + 1 and np.all(np.isfinite(d["Energy"]))


# ---- host-side files of the drivers' path: readFlowFile.m, legacy/writeFlowFile.m, legacy/flowToColor.m (+ maxFlow) -----------
def test_host_io_against_executed_source(pkg, O, tmp_path):
    d = np.load(os.path.join(GOLD, "refsrc_host_io.npz"))
    flow = d["flow"]
    fn = str(tmp_path / "a.flo")
    pkg.writeFlowFile(flow, fn)                                          # product writer == bytes legacy/writeFlowFile.m produced
    assert open(fn, "rb").read() == d["flo_bytes"].tobytes()
    assert np.array_equal(pkg.readFlowFile(fn), d["flo_read_back"]) and np.array_equal(d["flo_read_back"], flow)
    for tag in ("mf4", "mf05", "mfneg"):                                 # flowToColor(flow, maxFlow): oracle and C ABI, bit for bit
        mf = float(d[tag + "_maxflow"])
        for impl in (lambda: O.flow_to_color(flow, mf), lambda: pkg.flowToColor_mex(flow, mf)):
            r = impl()
            assert np.array_equal(r[0], d[tag + "_img"]) and np.array_equal(r[1], d[tag + "_flo"]), tag
            assert tuple(float(x) for x in r[2:6]) == tuple(d[tag + "_range"]) and np.array_equal(np.asarray(r[6], bool), d[tag + "_unknown"])
    assert not np.array_equal(d["mf4_img"], d["mf05_img"]) and (d["mf05_img"].astype(int).sum() < d["mf4_img"].astype(int).sum())


def test_host_io_live(pkg, O, tmp_path):
    sys.path.insert(0, GOLD)
    import make_refsrc_golden as G
    if not os.path.isdir(G.REF):
        pytest.skip("reference tree not present on this box (the committed vectors cover it)")
    from oracle.mlab.minimat import Interp
    from oracle.refbin import refbin
    interp = Interp([G.REF, os.path.join(G.REF, "legacy")])
    for seq in ("Venus", "rubberwhale"):
        fn = os.path.join(G.REF, "middlebury", seq, "flow10.flo")
        ref = interp.call("readFlowFile", fn)                            # readFlowFile.m executed on a shipped ground-truth file
        assert np.array_equal(pkg.readFlowFile(fn), ref)
        out = str(tmp_path / (seq + ".flo"))
        interp.call("writeFlowFile", ref, out, nargout=0)                # legacy/writeFlowFile.m reproduces the shipped file byte for byte
        assert open(out, "rb").read() == open(fn, "rb").read()
        src = interp.call("flowToColor", ref, nargout=7)                 # the .m source == the compiled binary == oracle == product
        for other in (refbin.flowToColor_mex(ref), O.flow_to_color(ref), pkg.flowToColor_mex(ref)):
            assert all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(src, other)), seq


def test_reference_rejects_row_vector_flow_live(pkg, O):
    """A finding, kept executable: the reference fails on a 1 x N flow field (N > 1) -- computeColor.m:57 indexes a column vector with
    a row vector and gets a column back, so the arithmetic after it no longer conforms.  The compiled binary reports it through
    Coder's run-time check, the .m source through MATLAB's assignment rule; the oracle and the product treat 1 x N like any field."""
    sys.path.insert(0, GOLD)
    import make_refsrc_golden as G
    if not os.path.isdir(G.REF):
        pytest.skip("reference tree not present on this box")
    from oracle.mlab.minimat import Interp, MatlabError
    from oracle.refbin import refbin
    flow = np.asfortranarray(np.random.default_rng(0).normal(0, 1, (1, 7, 2)))
    with pytest.raises(RuntimeError):
        refbin.flowToColor_mex(flow)
    with pytest.raises(MatlabError):
        Interp([G.REF, os.path.join(G.REF, "legacy")]).call("flowToColor", flow, nargout=7)
    a, b = O.flow_to_color(flow), pkg.flowToColor_mex(flow)
    assert a[0].shape == (1, 7, 3) and np.array_equal(a[0], b[0])
    col = np.asfortranarray(flow.transpose(1, 0, 2))                      # the transposed field is accepted by everyone and gives the same colours
    assert np.array_equal(refbin.flowToColor_mex(col)[0].transpose(1, 0, 2), a[0])


# ---- GPU: the CUDA path against the executed source ---------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", GPU_CASES)
def test_cuda_step_against_executed_source(pkg, O, name):
    from test_gpu_parity import _assert_state_close, _round_state
    d = load(name)
    cfg = config(O, d)
    VV = O.get_vv(d["I2"])
    variant = "super" if cfg.super else "full"
    before = _round_state(initial_state(O, d, cfg))           # the device keeps beliefs in fp32 (relative 6e-8 off the executed start)
    ref = probe_state(O, d, cfg, 1, d["T"])
    with pkg.Solver(options_from_cfg(cfg, T=d["T"]), d["I1"], d["I2"], variant=variant) as s:
        s.set_state(state_dict(before), T=d["T"], it=1, alpha=before.alpha)
        r = s.step(1)
        got = s.get_state()
    assert abs(r["Energy"][0] / d["Energy"][0] - 1) < 1e-5, (r["Energy"][0], d["Energy"][0])
    assert abs(r["ptdmu"][0] / float(d["p1_ptdmu"]) - 1) < 1e-3 and abs(r["ptdsigma"][0] / float(d["p1_ptdsigma"]) - 1) < 1e-3
    _assert_state_close(O, cfg, d["I1"], VV, got, ref, before, cfg.step0 / (1 + 1 / cfg.step_tau), where=name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", LONG)
def test_cuda_alpha_and_anneal_against_executed_source(pkg, O, name):
    from test_gpu_parity import _round_state
    d = load(name)
    cfg = config(O, d)
    variant = "super" if cfg.super else "full"
    probes = [int(k) for k in d["probes"]]
    with pkg.Solver(options_from_cfg(cfg, T=d["T"]), d["I1"], d["I2"], variant=variant) as s:
        for k in probes:
            if k + 1 not in probes or k < 400:
                continue
            T_next = float(d["p%d_T" % (k + 1)])
            before = _round_state(probe_state(O, d, cfg, k, T_next))
            s.set_state(state_dict(before), T=T_next, it=k + 1, alpha=before.alpha)
            r = s.step(1)
            got = s.get_state()
            assert abs(r["Energy"][0] / float(d["p%d_Energy" % (k + 1)]) - 1) < 1e-5, (k, r["Energy"][0])
            da = np.abs(np.ravel(d["p%d_alpha" % (k + 1)]) - before.alpha).max()
            assert np.abs(got["alpha"] - np.ravel(d["p%d_alpha" % (k + 1)])).max() < 1e-12 + 1e-3 * da, k
            if k + 2 in probes:
                assert abs(got["T"] - float(d["p%d_T" % (k + 2)])) < 1e-15, (k, got["T"])


# (last in the file and in the suite: written after the round's GPU budget was spent -- its CPU twin above and its script logic, run against an
# oracle-backed stand-in for the Solver, are verified; the thresholds are those the full-size late-state tests hold on B200)
@pytest.mark.gpu
def test_cuda_single_step_from_late_state_on_reference_data(pkg, O):
    from test_gpu_parity import _round_state
    from test_gpu_full_size import _assert_step_close
    d, cfg, I1, I2, state = _long_case(O, "grove2_window_L3K5_it16")
    k = REAL_LATE["grove2_window_L3K5_it16"][6]
    VV = O.get_vv(I2)
    before = _round_state(state(k))
    ref = before.copy()
    _, _, _, E, dm, ds = O.run(cfg, I1, VV, ref, k + 1, 10 ** 6, 1)           # the oracle from the fp32-rounded state, whole state
    with pkg.Solver(options_from_cfg(cfg), I1, I2) as s:
        s.set_state(state_dict(before), it=k + 1, alpha=before.alpha)
        r = s.step(1)
        got = s.get_state()
    print("late state it=%d: Energy rel %.2e (executed source), %.2e (oracle, same fp32 start); ptdmu rel %.2e" % (
        k + 1, r["Energy"][0] / d["Energy"][k] - 1, r["Energy"][0] / E[0] - 1, r["ptdmu"][0] / dm[0] - 1))
    assert abs(r["Energy"][0] / E[0] - 1) < 1e-5 and abs(r["Energy"][0] / d["Energy"][k] - 1) < 1e-4        # north_star: 1e-4
    # mean|G| (:69-70): 1e-4 relative plus the fp32 floor of tests/test_gpu_full_size.py -- this state has sigmas at the 0.01 floor and
    # correlations at the clamp, where rounding the INPUT to fp32 alone moves mean|dmu| by 1.7e-5 relative (measured with the oracle)
    prn = (1 - before.pn ** 2)[1:-1, 1:-1]
    floor_u = 3e-5 * float(np.mean(1.0 / (before.sigu[1:-1, 1:-1] * prn)))
    assert abs(r["ptdmu"][0] - dm[0]) < 1e-4 * dm[0] + floor_u and abs(r["ptdsigma"][0] - ds[0]) < 1e-4 * ds[0] + floor_u
    worst = _assert_step_close(got, ref, before, cfg.step0 / (1 + (k + 1) / cfg.step_tau), where="Grove2 window, it=%d" % (k + 1), max_bad_frac=1e-3)
    print("worst error / bound per field:", worst)
