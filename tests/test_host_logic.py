"""Host-side logic of the product that needs no GPU, checked against the oracle / known answers."""
import os
import struct

import numpy as np
import pytest


def test_gauss_hermite_product_vs_oracle(pkg, O):
    for n in range(1, 33):
        x, w = pkg.GaussHermite_2(n)
        assert x.shape == (n, 1) and w.shape == (n, 1)                 # column vectors, as GaussHermite_2.m:30-32
        if n >= 2:
            xo, wo = O.gauss_hermite(n)
            assert np.abs(x.ravel() - xo).max() < 1e-13 and np.abs(w.ravel() - wo).max() < 1e-14
        assert abs(w.sum() - np.sqrt(np.pi)) < 1e-13
        assert np.array_equal(x.ravel(), -x.ravel()[::-1])             # exactly symmetric tables
    with pytest.raises(pkg.QgmapError):
        pkg.GaussHermite_2(33)


def test_projsplx_product_vs_oracle(pkg, O):
    rng = np.random.default_rng(0)
    for m in (1, 2, 3, 5, 10):
        for _ in range(40):
            y = rng.normal(0, 1.5, m)
            assert np.allclose(pkg.projsplx(y), O.projsplx(y), atol=1e-15)


def test_flow_to_color_product_vs_oracle(pkg, O):
    rng = np.random.default_rng(1)
    flow = rng.normal(0, 3, (33, 47, 2))
    flow[rng.random((33, 47)) < 0.05] = 1e10
    flow[5, 5] = (0.0, 0.0)
    for mf in (None, 4.0):
        a = pkg.flowToColor_mex(flow, mf)
        b = O.flow_to_color(flow, -1.0 if mf is None else mf)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2:6] == b[2:6] and np.array_equal(a[6], b[6])
    assert a[0].dtype == np.uint8 and a[0].shape == (33, 47, 3) and a[6].dtype == bool


def test_read_flow_file(pkg, tmp_path):
    h, w = 5, 7
    u = np.arange(h * w, dtype=np.float32).reshape(h, w)
    v = -u
    inter = np.stack([u, v], axis=2).reshape(h, w * 2)
    p = tmp_path / "t.flo"
    with open(p, "wb") as f:
        f.write(struct.pack("<f", 202021.25) + struct.pack("<ii", w, h) + inter.tobytes())
    img = pkg.readFlowFile(str(p))
    assert img.shape == (h, w, 2) and np.array_equal(img[:, :, 0], u) and np.array_equal(img[:, :, 1], v)
    with pytest.raises(ValueError):
        pkg.readFlowFile(str(tmp_path / "t.txt"))                      # readFlowFile.m:47-49
    bad = tmp_path / "bad.flo"
    bad.write_bytes(struct.pack("<f", 1.0) + struct.pack("<ii", w, h))
    with pytest.raises(ValueError):
        pkg.readFlowFile(str(bad))                                     # wrong tag, readFlowFile.m:62-64


def test_rgb2gray_matlab_coefficients(pkg):
    rgb = np.array([[[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 255], [10, 20, 30]]], dtype=np.uint8)
    assert pkg.rgb2gray(rgb).tolist() == [[76, 150, 29, 255, 18]]      # MATLAB rgb2gray on uint8


def test_synthetic_pair_deterministic_and_consistent(pkg):
    a = pkg.synthetic_pair(48, 64, seed=5)
    b = pkg.synthetic_pair(48, 64, seed=5)
    assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3] == b[3]
    I1, I2, flow, (minu, maxu, minv, maxv) = a
    assert I1.min() == 0.0 and I1.max() == 255.0 and I1.flags.f_contiguous
    assert abs(maxu - 4.0) < 0.02 and abs(minu + 2.0) < 0.02 and abs(maxv - 2.0) < 1e-12
    # brightness constancy holds along the flow: I2(row+v, col+u) ~ I1(row, col) in the interior
    from importlib import import_module
    fr = import_module("gqmap-opticalflow_b200.frames")
    rows = np.arange(48.0).reshape(48, 1); cols = np.arange(64.0).reshape(1, 64)
    warped = fr._bicubic(I2, rows + flow[:, :, 1], cols + flow[:, :, 0])
    assert np.abs(warped - I1)[8:-8, 8:-8].mean() < 2.0


def test_make_config_maps_options(pkg):
    o = dict(K=9, L=3, temperature=0.2, drate=0.75, epsn=1e-6, lambdad=1, lambdas=16, minu=-2, maxu=3, minv=-1, maxv=1,
             alpha_mode="projsplx", sigma_max=17.0, alpha_start=700)
    c = pkg.make_config(o, 1)
    assert (c.K, c.L, c.temperature, c.drate, c.lambdas, c.minu, c.maxu, c.alpha_mode, c.sigma_max, c.alpha_start) == \
           (9, 3, 0.2, 0.75, 16.0, -2.0, 3.0, 1, 17.0, 700)
    assert c.anneal_every == 500 and c.step0 == 0.001                  # super-pixel constants kept
    with pytest.raises(ValueError):
        pkg.make_config(dict(o, alpha_mode="simplex"), 0)


def test_mex_gateways_compile_against_stub():
    mexdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gqmap-opticalflow_b200", "mex")
    import subprocess
    subprocess.check_call(["make", "-C", mexdir, "-B"], stdout=subprocess.DEVNULL)
    for f in ("gqmap_mex.o", "get_map_mex.o", "flowToColor_mex.o"):
        assert os.path.exists(os.path.join(mexdir, f))


def test_png_writer_round_trips(pkg, tmp_path):
    """imwrite replacement (gqmap_gpu_mixture.m:62): the PNG decodes (PIL) to exactly the image flowToColor_mex produced."""
    from PIL import Image
    rng = np.random.default_rng(3)
    for shape in ((1, 1), (7, 300), (388, 584)):                       # last one needs several stored deflate blocks
        img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        p = str(tmp_path / ("t%dx%d.png" % shape))
        pkg.imwrite(img, p)
        back = np.asarray(Image.open(p))
        assert back.shape == shape + (3,) and np.array_equal(back, img)
    with pytest.raises(pkg.QgmapError):
        pkg.imwrite(img, str(tmp_path / "no_such_dir" / "x.png"))
    with pytest.raises(ValueError):
        pkg.imwrite(img[:, :, 0], str(tmp_path / "y.png"))


def test_flo_files_round_trip(pkg, tmp_path):
    """readFlowFile.m / legacy/writeFlowFile.m: Python mirror and C ABI write identical files and read each other's."""
    import ctypes as C
    rng = np.random.default_rng(4)
    flow = np.asfortranarray(rng.normal(0, 5, (13, 29, 2)).astype(np.float32).astype(np.float64))
    flow[3, 4] = 1.6e9                                                  # unknown-flow marker survives (float32)
    a, b = str(tmp_path / "a.flo"), str(tmp_path / "b.flo")
    pkg.writeFlowFile(flow, a)
    lib = pkg._lib.lib
    assert lib.qgmap_write_flo(b.encode(), pkg._lib.dptr(flow), 13, 29) == 0
    assert open(a, "rb").read() == open(b, "rb").read()
    assert open(a, "rb").read(4) == b"PIEH" and struct.unpack("<f", b"PIEH")[0] == 202021.25
    assert np.array_equal(pkg.readFlowFile(b), flow)
    H, W = C.c_int(), C.c_int()
    assert lib.qgmap_read_flo(a.encode(), C.byref(H), C.byref(W), None) == 0 and (H.value, W.value) == (13, 29)
    out = np.zeros((13, 29, 2), order="F")
    assert lib.qgmap_read_flo(a.encode(), C.byref(H), C.byref(W), pkg._lib.dptr(out)) == 0 and np.array_equal(out, flow)
    assert lib.qgmap_read_flo(str(tmp_path / "a.txt").encode(), C.byref(H), C.byref(W), None) != 0       # extension check
    open(str(tmp_path / "bad.flo"), "wb").write(b"XXXX" + bytes(8))
    assert lib.qgmap_read_flo(str(tmp_path / "bad.flo").encode(), C.byref(H), C.byref(W), None) != 0     # wrong tag
    with pytest.raises(ValueError):
        pkg.writeFlowFile(flow[:, :, :1], a)
    with pytest.raises(ValueError):
        pkg.writeFlowFile(flow, str(tmp_path / "a.dat"))


def test_save_results_schema(pkg, tmp_path):
    """optical_flow.m:28: the .mat carries options, AEPE, mu, sigma, alpha, Energy, logP."""
    from scipy.io import loadmat
    p = str(tmp_path / "r.mat")
    opts = dict(K=3, L=2, its=5, dir="x", lambdas=5.0)
    pkg.save_results(p, opts, np.zeros((4, 5, 2, 2)), np.ones((4, 5, 2, 2)), np.full((1, 1, 2), 0.5), np.zeros((5, 1)), np.zeros((5, 1)), np.zeros((5, 1)))
    m = loadmat(p)
    assert {"options", "AEPE", "mu", "sigma", "alpha", "Energy", "logP"} <= set(m)
    assert m["mu"].shape == (4, 5, 2, 2) and m["alpha"].shape == (1, 1, 2) and int(m["options"]["K"][0, 0][0, 0]) == 3


def test_ctf_host_glue(pkg):
    """legacy/optical_flow_ctf.m:22-30 host glue: imresize (bicubic, antialiased), interp2 (linear, NaN outside), fillmissing."""
    from PIL import Image
    c = pkg.ctf
    rng = np.random.default_rng(0)
    A = rng.random((48, 64)) * 255
    assert np.allclose(c.imresize(A, 1.0), A, atol=1e-12)
    for sc in (0.5, 0.25, 2.0):
        assert np.allclose(c.imresize(np.full((20, 30), 7.0), sc), 7.0, atol=1e-12)                # weights are normalised
        B = c.imresize(A, sc)
        assert B.shape == (int(np.ceil(48 * sc)), int(np.ceil(64 * sc)))
        # independent implementation of the same antialiased Keys(-0.5) resampling (PIL, float32): equal away from the border
        P = np.asarray(Image.fromarray(A.astype(np.float32), mode="F").resize((B.shape[1], B.shape[0]), Image.BICUBIC), dtype=np.float64)
        assert np.abs(B - P)[3:-3, 3:-3].max() < 1e-3
    R = np.add.outer(np.arange(20.0), 2 * np.arange(30.0))
    U = c.imresize(R, 2)                                                                            # cubic convolution reproduces ramps
    assert abs((U[10, 11] - U[10, 10]) - 1.0) < 1e-12 and abs((U[11, 10] - U[10, 10]) - 0.5) < 1e-12
    assert c.imresize(np.zeros((33, 47, 2)), size=(17, 24)).shape == (17, 24, 2)
    V = np.arange(12.0).reshape(3, 4)
    q = c.interp2_linear(V, np.array([[1.5, 4.0, 4.1, 0.99]]), np.array([[1.0, 3.0, 2.0, 1.0]]))
    assert q[0, 0] == 0.5 and q[0, 1] == 11.0 and np.isnan(q[0, 2]) and np.isnan(q[0, 3])
    F = np.array([[1, np.nan, np.nan, np.nan, 5.0], [np.nan] * 5, [np.nan, 2, np.nan, np.nan, np.nan]])
    assert np.array_equal(c.fillmissing_nearest(F, 1)[0], [1, 1, 5, 5, 5]) and np.isnan(c.fillmissing_nearest(F, 1)[1]).all()
    G = c.fillmissing_nearest(c.fillmissing_nearest(F, 0), 1)
    assert not np.isnan(G).any() and np.array_equal(G[1], [1, 2, 2, 5, 5])
    o = c.ctf_options(dict(K=11, its=3000, epsn=1e-6, lambdas=5, lambdad=1), np.dstack([np.full((4, 4), -2.0), np.full((4, 4), 3.0)]))
    assert (o["L"], o["step0"], o["sigma_step_scale"], o["sigma_max"], o["corr_tor"], o["minu"], o["maxv"]) == (1, 0.07, 0.3, 25.0, 0.999, -2.0, 3.0)


def test_ctf_pyramid_glue_is_consistent(pkg, monkeypatch):
    """The pyramid loop of legacy/optical_flow_ctf.m:21-36 with the per-level solver replaced by one that returns the exact
    residual flow: resizing, x2 scaling, warp sign (I1_w(x) = I1(x - warp)) and accumulation must then reproduce the ground truth,
    and the warped first frame must match the second frame at every level (no GPU involved)."""
    c = pkg.ctf
    M, N = 96, 128
    I1, I2, flow, _ = pkg.synthetic_pair(M, N, seed=3, flow_scale=2.0)
    seen = []

    def perfect(options, I1w, I2l, GRDT, seed=0, device=-1, aepe_target=None):
        m, n = I1w.shape
        seen.append((float(np.abs(I1w - I2l)[4:-4, 4:-4].mean()), float(np.abs(aepe_target).max())))
        assert GRDT.shape == (M, N, 2)                                                     # :31 passes trueFlow.*scale, full size
        return aepe_target.copy(), np.ones((m, n, 2)), np.zeros((m, n, 2, 2)), np.full(3, np.nan), np.ones(3)
    monkeypatch.setattr(c, "gqmap_ctf", perfect)
    warp, levels = c.optical_flow_ctf(I1, I2, flow, dict(K=3, its=3), scales=(1 / 4, 1 / 2, 1))
    assert np.abs(warp - flow).max() < 1e-9 and all(lv["aepe_after"] < 1e-9 for lv in levels)
    # level 1 starts from zero warp (photometric error of the raw pair), later levels from the up-scaled previous flow: the warped
    # frame is then close to the second frame (what remains is the blur of the bilinear warp on this fine texture) and the
    # residual flow is a fraction of a pixel
    assert seen[1][0] < 0.6 * seen[0][0] and seen[2][0] < 0.6 * seen[0][0], seen
    assert seen[1][1] < 0.5 and seen[2][1] < 0.5, seen
    x, y = np.meshgrid(np.arange(1, N + 1.0), np.arange(1, M + 1.0))
    right = np.nanmean(np.abs(c.interp2_linear(I1, x - flow[:, :, 0], y - flow[:, :, 1]) - I2)[6:-6, 6:-6])
    wrong = np.nanmean(np.abs(c.interp2_linear(I1, x + flow[:, :, 0], y + flow[:, :, 1]) - I2)[6:-6, 6:-6])
    assert right < 0.4 * wrong                                                              # sign of the warp (:28-29)
