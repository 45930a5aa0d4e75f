"""the per-level solver of the reference's coarse-to-fine driver, legacy/gqmap_ctf.m, EXECUTED unmodified under
the interpreter (tests/golden/make_ctf_golden.py -> refsrc_ctf.npz) against the oracle run with gqmap_ctf's constants
(legacy/gqmap_ctf.m:5,15-16,27,34-37: corr_tor 0.999, sigma = rand+3, constant step 0.07, sigma step x0.3, sigma <= 25, one Gaussian,
no entropy term) and its data term, a nearest lookup into the 64x upsampled second frame (:10,:96; oracle.set_nearest_lookup).
The product's gqmap_ctf keeps the live solver's exact bicubic sample instead (include/qgmap.h); the last test
measures what that changes."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _problem(O):
    d = np.load(os.path.join(HERE, "golden", "refsrc_ctf.npz"))
    Mo, No, K, its = (int(v) for v in d["meta"])
    g = d["GRDT"]
    minu, maxu, minv, maxv = g[:, :, 0].min(), g[:, :, 0].max(), g[:, :, 1].min(), g[:, :, 1].max()       # legacy/gqmap_ctf.m:4
    cfg = O.make_config(Mo, No, 1, K, lambdas=5.0, epsn=1e-6, minu=minu, maxu=maxu, minv=minv, maxv=maxv, corr_tor=0.999, step0=0.07,
                        step_tau=1e300, sigma_step_scale=0.3, sigma_max=25.0, alpha_start=10 ** 9)
    sh = (Mo, No, 1)
    st = O.State((minu + d["draw0"] * (maxu - minu)).reshape(sh, order="F"), (minv + d["draw1"] * (maxv - minv)).reshape(sh, order="F"),
                 (d["draw2"] + 3).reshape(sh, order="F"), (d["draw3"] + 3).reshape(sh, order="F"), np.zeros(sh), np.zeros(sh + (2, 2)),
                 np.zeros(1), alpha=np.ones(1))                                                           # :13-18
    return d, cfg, st, its


def test_oracle_reproduces_executed_gqmap_ctf(O):
    d, cfg, st, its = _problem(O)
    I1, I2 = d["I1"], d["I2"]
    VV = O.get_vv(I2)
    O.set_nearest_lookup(O.interp2_cubic_refine(I2, 6), 6)                                                # :10  rfc = 6
    try:
        for it in range(1, its + 1):
            n, _, stopped, E, dm, ds = O.run(cfg, I1, VV, st, it, its, 1)
            assert n == 1
            assert abs(np.pi * E[0] / d["Energy"][it - 1] - 1) < 1e-12          # :39 sums fval without the 1/pi of the mixture solver
            assert abs(dm[0] / float(d["p%d_ptdmu" % it]) - 1) < 1e-10 and abs(ds[0] / float(d["p%d_ptdsigma" % it]) - 1) < 1e-10   # :46
            for f, a in (("muu", st.muu), ("muv", st.muv), ("sigmau", st.sigu), ("sigmav", st.sigv), ("pn", st.pn)):
                assert np.abs(a[:, :, 0] - d["p%d_%s" % (it, f)]).max() < 1e-9, (it, f)
            assert np.abs(st.rou[:, :, 0] - d["p%d_rou" % it]).max() < 1e-9
            aepe = np.sqrt((d["GRDT"][1:-1, 1:-1, 0] - st.muu[1:-1, 1:-1, 0]) ** 2 + (d["GRDT"][1:-1, 1:-1, 1] - st.muv[1:-1, 1:-1, 0]) ** 2).mean()
            assert abs(aepe - d["AEPE"][it - 1]) < 1e-12                                                  # :38
        assert stopped                                                                                    # :49  it > its
        assert np.abs(np.dstack([st.muu[:, :, 0], st.muv[:, :, 0]]) - d["mu"]).max() < 1e-9               # :152-154 outputs
        assert np.abs(np.dstack([st.sigu[:, :, 0], st.sigv[:, :, 0]]) - d["sigma"]).max() < 1e-9
    finally:
        O.set_nearest_lookup(None)


def test_interp2_restatement_properties(O):
    """interp2(V,k,'cubic') as restated (third party, unpinned): it interpolates the samples, reproduces quadratics exactly (Keys
    a = -0.5 with the 3a-3b+c ring) and has the refined size (M-1)2^k+1."""
    rng = np.random.default_rng(3)
    V = np.asfortranarray(rng.random((7, 9)))
    R = O.interp2_cubic_refine(V, 3)
    assert R.shape == (49, 65) and np.array_equal(R[::8, ::8], V)
    yy, xx = np.mgrid[0:7, 0:9].astype(np.float64)
    Q = np.asfortranarray(0.3 * xx ** 2 - 0.2 * xx * yy + 0.1 * yy ** 2 + xx - 2 * yy + 4)
    y2, x2 = np.mgrid[0:49, 0:65] / 8.0
    assert np.abs(O.interp2_cubic_refine(Q, 3) - (0.3 * x2 ** 2 - 0.2 * x2 * y2 + 0.1 * y2 ** 2 + x2 - 2 * y2 + 4)).max() < 1e-12


def test_nearest_lookup_vs_exact_bicubic(O):
    """What the product's deliberate difference costs: gqmap_ctf.m quantises the sample position to 1/64 px (nearest lookup);
    the product samples the same interpolant exactly.  After the golden run's three iterations the two data terms leave the means
    within a few hundredths of a pixel of each other -- far below what the next pyramid level's re-initialisation over the whole
    clamp range (:13-14) discards anyway."""
    d, cfg, st, its = _problem(O)
    I1, I2 = d["I1"], d["I2"]
    VV = O.get_vv(I2)
    exact = st.copy()
    O.run(cfg, I1, VV, exact, 1, its, its)
    O.set_nearest_lookup(O.interp2_cubic_refine(I2, 6), 6)
    try:
        O.run(cfg, I1, VV, st, 1, its, its)
    finally:
        O.set_nearest_lookup(None)
    dmu = max(np.abs(st.muu - exact.muu).max(), np.abs(st.muv - exact.muv).max())
    assert 0 < dmu < 0.05 and np.abs(st.sigu - exact.sigu).max() < 0.05, dmu
