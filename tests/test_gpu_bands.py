"""Row-band decomposition on ONE GPU: nbands handles exchange halo rows and global sums through the same code
path a multi-GPU single-process host uses (qgmap_group_*).  The N-band result must equal the 1-band result bit for bit in
the beliefs (identical fp32 arithmetic per pixel) and to fp64 rounding in the reductions."""
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


@pytest.mark.parametrize("variant,shape,L,K,T,nbands", [
    ("full", (61, 70), 2, 3, 0.0, 2), ("full", (61, 70), 2, 3, 0.0, 3), ("full", (96, 45), 3, 5, 0.2, 4),
    ("super", (96, 128), 2, 3, 0.2, 2), ("super", (128, 96), 3, 5, 0.2, 5),
])
def test_bands_match_single_domain(pkg, O, variant, shape, L, K, T, nbands):
    sup = variant == "super"
    Mo, No = shape
    cfg, I1, I2, st = make_problem(O, Mo, No, L, K, super=sup, seed=5, T=T, small_sigma=True)
    opts = options_from_cfg(cfg, T=T, alpha_scale=1e-5)
    n = 14
    with pkg.Solver(opts, I1, I2, variant=variant) as s:
        s.set_state(state_dict(st), T=T, it=495)            # crosses it=500: alpha update and (super) anneal are exercised
        r1 = s.step(n)
        a = s.get_state()
    with pkg.BandGroup(opts, I1, I2, nbands, variant=variant) as g:
        assert g.nbands == nbands
        g.set_state(state_dict(st), T=T, it=495)
        rb = g.step(n)
        b = g.get_state()
    assert r1["n_done"] == rb["n_done"] == n
    for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou"):
        assert np.array_equal(a[f], b[f]), f                 # bit-identical beliefs, halo rows included
    assert np.abs(rb["Energy"] / r1["Energy"] - 1).max() < 1e-12
    assert np.abs(rb["ptdmu"] / r1["ptdmu"] - 1).max() < 1e-12
    assert np.abs(a["alpha"] - b["alpha"]).max() < 1e-14 and a["it"] == b["it"] == 495 + n
    assert a["T"] == b["T"]


def test_band_group_stop_rule(pkg, O):
    cfg, I1, I2, st = make_problem(O, 48, 40, 1, 3, seed=8)
    with pkg.BandGroup(options_from_cfg(cfg), I1, I2, 3) as g:
        g.set_state(state_dict(st))
        r = g.step(10, its=6)
        assert (r["n_done"], r["stopped"]) == (6, True)
        assert g.step(3, its=6)["n_done"] == 0


@pytest.mark.parametrize("variant", ["full", "super"])
def test_solve_with_options_devices_matches_single_domain(pkg, O, variant, tmp_path):
    """gqmap_gpu_mixture(options,I1,I2) with the new optional options.devices (one row band per entry; here three bands that share
    the GPU) returns the same beliefs bit for bit and the same AEPE / Energy / logP histories as the undivided call, PNG dumps
    included."""
    sup = variant == "super"
    Mo, No = (96, 128) if sup else (60, 72)
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(Mo, No)
    opts = dict(K=3, L=2, its=13, temperature=0.2 if sup else 0.0, drate=0.75, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu,
                minv=minv, maxv=maxv, seed=5, log_every=4, trueFlow=flow, unknownIdx=np.zeros((Mo, No), bool), alpha_start=2, alpha_scale=1e-5)
    fn = pkg.gqmap_gpuSuper_mix_entropy if sup else pkg.gqmap_gpu_mixture
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    a = fn(dict(opts, dir=str(tmp_path / "a")), I1, I2)
    b = fn(dict(opts, dir=str(tmp_path / "b"), devices=[0, 0, 0]), I1, I2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])                       # mu, sigma
    assert np.abs(a[2] - b[2]).max() < 1e-14                                               # alpha
    for k in (3, 4, 5):                                                                    # AEPE, Energy, logP (NaN pattern included)
        assert np.array_equal(np.isnan(a[k]), np.isnan(b[k]))
        m = ~np.isnan(a[k])
        assert np.abs(b[k][m] / a[k][m] - 1).max() < 1e-11, k
    assert sorted(p.name for p in (tmp_path / "b").iterdir()) == ["1.png", "12.png", "4.png", "8.png"]
    for p in (tmp_path / "a").iterdir():
        assert p.read_bytes() == (tmp_path / "b" / p.name).read_bytes()


@pytest.mark.parametrize("variant,shape,L,K,T,nbands,form", [
    ("full", (61, 70), 2, 3, 0.0, 2, "tile"), ("full", (96, 45), 3, 5, 0.2, 3, "tile"), ("full", (96, 45), 3, 5, 0.2, 3, "walk"),
    ("super", (128, 96), 3, 5, 0.2, 2, "tile"),
])
def test_peer_memory_transport_on_one_gpu(pkg, O, monkeypatch, variant, shape, L, K, T, nbands, form):
    """The peer-memory exchange (csrc/qgmap_peer.cuh: boundary rows stored into the neighbour's halo rows, sums and flags through
    the mailboxes, waits inside the kernels) on a 1-GPU box: the bands' streams run concurrently on the same device
    (QGMAP_GROUP_TRANSPORT=p2p-shared, a test-only switch).  One lane per belief = the exchange fused into the iteration kernel
    (tiled and row-walking form), CUDA-graph launches included (30 iterations > one 25-node graph); four lanes per belief =
    iteration kernel + publish kernel.  Same bits as the single domain, across two step calls (generation tags)."""
    monkeypatch.setenv("QGMAP_GROUP_TRANSPORT", "p2p-shared")
    monkeypatch.setenv("QGMAP_ITER", form)
    sup = variant == "super"
    Mo, No = shape
    cfg, I1, I2, st = make_problem(O, Mo, No, L, K, super=sup, seed=5, T=T, small_sigma=True)
    opts = options_from_cfg(cfg, T=T, alpha_scale=1e-5)
    with pkg.Solver(opts, I1, I2, variant=variant) as s:
        s.set_state(state_dict(st), T=T, it=495)
        r1 = s.step(36)
        a = s.get_state()
    with pkg.BandGroup(opts, I1, I2, nbands, variant=variant) as g:
        g.set_state(state_dict(st), T=T, it=495)
        ra = g.step(6)
        rb = g.step(30)
        b = g.get_state()
    assert ra["n_done"] == 6 and rb["n_done"] == 30
    for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou"):
        assert np.array_equal(a[f], b[f]), f
    E = np.concatenate([ra["Energy"], rb["Energy"]])
    assert np.abs(E / r1["Energy"] - 1).max() < 1e-12 and np.abs(a["alpha"] - b["alpha"]).max() < 1e-14


class _Solo:
    """world-size-1 stand-in for torch.distributed (rendezvous is all the band solver asks of it)."""
    @staticmethod
    def get_rank():
        return 0

    @staticmethod
    def get_world_size():
        return 1


@pytest.mark.parametrize("variant", ["full", "super"])
def test_band_solver_equals_one_call_solver(pkg, variant):
    """dist.gqmap_gpu_mixture_bands -- the multi-process twin of gqmap_gpu_mixture(options.devices): the reference's loop with the
    monitoring block evaluated where the rows live (qgmap_monitor_partial) -- returns what the one-call solver returns: same
    beliefs bit for bit, same AEPE / Energy / logP histories incl. their NaN / zero prefill (gqmap_gpu_mixture.m:16,:52-68,:183-188)."""
    sup = variant == "super"
    Mo, No = (96, 128) if sup else (60, 72)
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(Mo, No)
    unk = np.zeros((Mo, No), bool)
    unk[5:9, 7:30] = True
    opts = dict(K=3, L=2, its=13, temperature=0.2 if sup else 0.0, drate=0.75, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu,
                minv=minv, maxv=maxv, seed=5, log_every=4, trueFlow=flow, unknownIdx=unk, alpha_start=2, alpha_scale=1e-5)
    fn = pkg.gqmap_gpuSuper_mix_entropy if sup else pkg.gqmap_gpu_mixture
    a = fn(opts, I1, I2)
    b = pkg.dist.gqmap_gpu_mixture_bands(opts, I1, I2, _Solo, variant=variant)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for k in (3, 4, 5):
        assert a[k].shape == b[k].shape and np.array_equal(np.isnan(a[k]), np.isnan(b[k]))
        m = ~np.isnan(a[k])
        assert np.abs(b[k][m] - a[k][m]).max() <= 1e-11 * np.abs(a[k][m]).max(), k
    assert (~np.isnan(a[3])).sum() == 4                        # it = 1, 4, 8, 12
