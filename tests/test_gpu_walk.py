"""The row-walking form of the full-resolution iteration kernel (csrc/qgmap_walk.cuh, selected with QGMAP_ITER=walk): a belief's
update must not depend on how the grid is cut into strips (options.strip_rows = rows one warp walks), nor on strips vs row bands
-- bit for bit, like the N-band == 1-band property of tests/test_gpu_bands.py.  The default tiled kernel evaluates the same
arithmetic (same device functions) with another reduction layout: the two agree to fp32 rounding."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ("muu", "muv", "sigmau", "sigmav", "pn", "rou")


def _run(pkg, opts, I1, I2, n, **extra):
    with pkg.Solver(dict(opts, **extra), I1, I2) as s:
        s.init_state(11)
        r = s.step(n)
        return s.get_state(), r


@pytest.mark.parametrize("L,K", [(2, 5), (3, 3), (1, 9), (2, 4)])
def test_strip_rows_invariance(pkg, monkeypatch, L, K):
    monkeypatch.setenv("QGMAP_ITER", "walk")
    Mo, No = 75, 100                        # 4 column strips, the last one ragged; 73 interior rows: ragged last strip for most heights
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(Mo, No, seed=5)
    opts = dict(K=K, L=L, temperature=0.05, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv,
                maxv=maxv, alpha_start=2, alpha_scale=1e-5)
    n = 8
    ref, rr = _run(pkg, opts, I1, I2, n, strip_rows=1)
    assert rr["n_done"] == n and np.isfinite(rr["Energy"]).all()
    for rows in (2, 3, 7, 16, 73, 200):
        got, rg = _run(pkg, opts, I1, I2, n, strip_rows=rows)
        for f in FIELDS:
            assert np.array_equal(ref[f], got[f]), (rows, f)
        assert np.abs(rg["Energy"] / rr["Energy"] - 1).max() < 1e-13 and np.abs(got["alpha"] - ref["alpha"]).max() < 1e-14
    with pkg.BandGroup(opts, I1, I2, 3) as g:                      # row bands cut the strips somewhere else again
        g.init_state(11)
        g.step(n)
        c = g.get_state()
    for f in FIELDS:
        assert np.array_equal(ref[f], c[f]), f


def test_walk_vs_tiled_kernel(pkg, monkeypatch):
    """Both kernels restate gqmap_gpu_mixture.m:27-50 with the same device functions: from the same state, a few steps agree to
    fp32 rounding of the gradients, and the reductions (fp32 row sums, fp64 across rows) to 1e-6."""
    Mo, No = 64, 96
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(Mo, No, seed=6)
    opts = dict(K=5, L=2, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
    b, rb = _run(pkg, opts, I1, I2, 1)
    monkeypatch.setenv("QGMAP_ITER", "walk")
    a, ra = _run(pkg, opts, I1, I2, 1, strip_rows=4)
    assert abs(ra["Energy"][0] / rb["Energy"][0] - 1) < 1e-6
    for f in ("muu", "muv", "sigmau", "sigmav"):
        assert np.abs(a[f] - b[f]).max() < 2e-4, f


def test_sample_paths_are_bit_identical(pkg, monkeypatch):
    """The full-resolution kernel picks a node-sample loop per warp: clamped (clouds that reach the image border), inside + tap
    cache (fp32 entries), inside + one-sector fp16 entries (wide beliefs on grey-level frames).  All evaluate
    gqmap_gpu_mixture.m:156-179 with the same operations in the same order, so moving the wide/cache boundary or switching the fp16
    layout off must not change a single bit -- from the wide random init and further along the ascent."""
    Mo, No = 120, 160
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(Mo, No, seed=9, grey_levels=True)
    opts = dict(K=5, L=2, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv,
                alpha_start=2, alpha_scale=1e-6)

    def run():
        with pkg.Solver(opts, I1, I2) as s:
            s.init_state(3)
            s.step(8)
            a = s.get_state()
            s.step(3000)
            s.step(8)
            return a, s.get_state()
    wide, late = run()
    assert np.median(wide["sigmau"]) > np.median(late["sigmau"])              # beliefs narrow along the way; the two extreme settings of
                                                                              # QGMAP_WIDE_REACH below force every warp through either loop
    for env in (dict(QGMAP_TAPS="f32"), dict(QGMAP_WIDE_REACH="0.0"), dict(QGMAP_WIDE_REACH="100")):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        w2, l2 = run()
        for k in env:
            monkeypatch.delenv(k)
        for f in FIELDS:
            assert np.array_equal(wide[f], w2[f]), (env, "wide", f)
            assert np.array_equal(late[f], l2[f]), (env, "late", f)
