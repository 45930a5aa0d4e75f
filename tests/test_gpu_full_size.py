"""BASELINE.json's full sizes, checked through size-independent properties (the oracle needs minutes per iteration batch there):
row-band invariance (N bands == one domain, bit for bit), run-to-run determinism, frozen borders, clamp ranges, and the exact
bookkeeping of the histories -- on 640x480 (configs[1], [2], [4] shapes) and on the 3840x2160 frame of configs[3]."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _opts(rng4, L, K, sup):
    minu, maxu, minv, maxv = rng4
    return dict(K=K, L=L, temperature=0.2 if sup else 0.0, drate=0.75, epsn=1e-6, lambdad=1.0, lambdas=16.0 if sup else 5.0,
                minu=minu, maxu=maxu, minv=minv, maxv=maxv, alpha_start=3, alpha_scale=1e-5)


@pytest.mark.parametrize("variant,L,K,nb", [("full", 3, 5, 3), ("full", 2, 9, 2), ("super", 3, 5, 4)])
def test_640x480_band_invariance_determinism_borders(pkg, variant, L, K, nb):
    sup = variant == "super"
    Mo, No = 480, 640
    I1, I2, flow, rng4 = pkg.synthetic_pair(Mo, No)
    opts = _opts(rng4, L, K, sup)
    n = 12
    runs = []
    for rep in range(2):
        with pkg.Solver(opts, I1, I2, variant=variant) as s:
            s.init_state(7)
            init = s.get_state()
            r = s.step(n)
            runs.append((s.get_state(), r))
    a, ra = runs[0]
    b, rb = runs[1]
    for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou"):
        assert np.array_equal(a[f], b[f]), f                                    # deterministic: fixed-order reductions, no atomics on data
        for sl in ((0,), (-1,), (slice(None), 0), (slice(None), -1)):           # frozen border rows / columns (:41-46 interior only)
            assert np.array_equal(a[f][sl], init[f][sl]), f
    assert np.array_equal(ra["Energy"], rb["Energy"]) and ra["n_done"] == n and np.isfinite(ra["Energy"]).all()
    assert a["muu"].min() >= np.float32(rng4[0]) - 1e-6 and a["muu"].max() <= np.float32(rng4[1]) + 1e-6      # :41 clamp
    assert a["sigmau"].min() >= np.float32(0.01) and np.abs(a["rou"]).max() <= 1 - 1e-5 + 1e-7               # :43,:45 clamps
    assert abs(a["alpha"].sum() - 1) < 1e-12 and not np.array_equal(a["alpha"], init["alpha"])               # :50 alpha moved, on the simplex
    with pkg.BandGroup(opts, I1, I2, nb, variant=variant) as g:
        g.init_state(7)
        rg = g.step(n)
        c = g.get_state()
    for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou"):
        assert np.array_equal(a[f], c[f]), f                                    # N bands == one domain, bit for bit
    assert np.abs(rg["Energy"] / ra["Energy"] - 1).max() < 1e-12 and np.abs(a["alpha"] - c["alpha"]).max() < 1e-14


def test_4k_band_invariance(pkg):
    """configs[3]: one 3840x2160 pair, L=3, K=5 -- two row bands against the undivided frame, through the one-call solver."""
    Mo, No = 2160, 3840
    I1, I2, flow, rng4 = pkg.synthetic_pair(Mo, No)
    opts = dict(_opts(rng4, 3, 5, False), its=6, seed=3, log_every=5)
    a = pkg.gqmap_gpu_mixture(opts, I1, I2)
    b = pkg.gqmap_gpu_mixture(dict(opts, devices=[0, 0]), I1, I2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.abs(a[2] - b[2]).max() < 1e-14
    assert np.abs(b[4] / a[4] - 1).max() < 1e-12                                # Energy history
    m = ~np.isnan(a[5])
    assert m.sum() == 2 and np.array_equal(np.isnan(a[5]), np.isnan(b[5])) and np.abs(b[5][m] / a[5][m] - 1).max() < 1e-11   # logP at it 1, 5
    assert a[0].shape == (Mo, No, 3, 2) and np.isfinite(a[0]).all()
