"""The C-ABI boundary: libqgmap.so loads, exports every symbol include/qgmap.h declares, the config struct matches, and
-- without a GPU -- every compute entry point fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qgmap.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qgmap_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    names = declared_symbols()
    assert len(names) >= 25
    lib = C.CDLL(pkg.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libqgmap.so does not export %s declared in include/qgmap.h" % n
    from_py = set(pkg._lib.SIGNATURES)
    assert from_py == set(names), (from_py ^ set(names))


def test_exports_are_c_linkage(pkg):
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for n in declared_symbols():
        assert n in exported, n


def test_config_struct_matches_header(pkg):
    cfg = pkg.QgmapConfig()
    for variant, (smax, step0, tau, anneal, lam) in ((0, (23.0, 0.1, 8000.0, 0, 5.0)), (1, (25.0, 0.001, 4000.0, 500, 16.0))):
        assert pkg._lib.lib.qgmap_config_defaults(C.byref(cfg), variant) == 0
        assert cfg.struct_size == C.sizeof(pkg.QgmapConfig)
        # constants hard-wired in gqmap_gpu_mixture.m:7,25,27,43,50,83 / gqmap_gpuSuper_mix_entropy.m:26,42,72
        assert (cfg.sigma_min, cfg.sigma_max, cfg.step0, cfg.step_tau, cfg.anneal_every, cfg.lambdas) == (0.01, smax, step0, tau, anneal, lam)
        assert (cfg.corr_tor, cfg.alpha_start, cfg.alpha_scale, cfg.tor, cfg.T_floor, cfg.log_every) == (1 - 1e-5, 500, 1e-7, 1e-4, 0.001, 300)
    assert pkg._lib.lib.qgmap_config_defaults(C.byref(cfg), 7) == -1
    assert pkg._lib.lib.qgmap_version() == 100


def test_bad_arguments_are_rejected(pkg):
    lib = pkg._lib.lib
    cfg = pkg.QgmapConfig()
    lib.qgmap_config_defaults(C.byref(cfg), 0)
    h = C.c_void_p()
    I = np.zeros((8, 8), order="F")
    p = I.ctypes.data_as(C.POINTER(C.c_double))
    cfg.L = 11                                          # get_map_mex limit: L <= 10
    assert lib.qgmap_create(C.byref(cfg), p, p, 8, 8, C.byref(h)) == -1 and b"L=11" in lib.qgmap_last_error(None)
    cfg.L, cfg.K = 1, 33
    assert lib.qgmap_create(C.byref(cfg), p, p, 8, 8, C.byref(h)) == -1
    cfg.K = 3
    cfg.struct_size = 8
    assert lib.qgmap_create(C.byref(cfg), p, p, 8, 8, C.byref(h)) == -1 and b"ABI" in lib.qgmap_last_error(None)
    lib.qgmap_config_defaults(C.byref(cfg), 1)
    assert lib.qgmap_create(C.byref(cfg), p, p, 10, 8, C.byref(h)) == -1       # super needs sides divisible by 4
    assert lib.qgmap_create(C.byref(cfg), None, p, 8, 8, C.byref(h)) == -1
    assert lib.qgmap_destroy(None) == -1 and lib.qgmap_step(None, 1, 1, None, None, None, None, None) == -1
    with pytest.raises(ValueError):
        pkg.get_map_mex([0.5, 0.5], np.zeros((4, 4, 2)), np.ones((4, 4, 3)), np.zeros((4, 4, 2)), np.ones((4, 4, 2)))
    with pytest.raises(ValueError):
        pkg.flowToColor_mex(np.zeros((4, 4, 3)))         # "image must have two bands"
    with pytest.raises(KeyError):
        pkg.make_config(dict(K=3, L=1), 0)               # options fields of gqmap_gpu_mixture.m:3-6 are required


def test_no_cpu_fallback_without_gpu(pkg):
    """On a box without a CUDA device every compute call must fail with QGMAP_ERR_CUDA (never compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device failure path cannot be exercised")
    with pytest.raises(pkg.QgmapError) as e:
        pkg.get_map_mex([1.0], np.zeros((4, 4, 1)), np.ones((4, 4, 1)), np.zeros((4, 4, 1)), np.ones((4, 4, 1)))
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)
    opts = dict(K=3, L=1, temperature=0, drate=0.5, epsn=1e-6, lambdad=1, lambdas=5, minu=-1, maxu=1, minv=-1, maxv=1, its=3)
    with pytest.raises(pkg.QgmapError) as e:
        pkg.gqmap_gpu_mixture(opts, np.zeros((8, 8)), np.zeros((8, 8)))
    assert e.value.status == -2
    with pytest.raises(pkg.QgmapError):
        pkg.Solver(opts, np.zeros((8, 8)), np.zeros((8, 8)))
    with pytest.raises(pkg.QgmapError) as e:                                   # options.devices -> qgmap_group_solve
        pkg.gqmap_gpu_mixture(dict(opts, devices=[0, 1]), np.zeros((8, 8)), np.zeros((8, 8)))
    assert e.value.status == -2
    with pytest.raises(pkg.QgmapError):
        pkg.BandGroup(opts, np.zeros((8, 8)), np.zeros((8, 8)), 2)
    with pytest.raises(pkg.QgmapError):                                        # coarse-to-fine driver: the solver is the CUDA one
        pkg.optical_flow_ctf(np.zeros((16, 16)), np.zeros((16, 16)), np.zeros((16, 16, 2)), dict(K=3, its=2, epsn=1e-6, lambdas=5, lambdad=1),
                             scales=(0.5, 1))
    lib = pkg._lib.lib
    assert lib.qgmap_band_p2p_export(None, None) == -1 and lib.qgmap_band_p2p_connect(None, 0, 2, None) == -1


def test_product_does_not_import_oracle():
    """The product path must not route through the oracle: no file of the package mentions it."""
    pk = os.path.join(ROOT, "gqmap-opticalflow_b200")
    for dp, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".m")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f in ("mex.h",), os.path.join(dp, f)


def test_header_is_plain_c_and_usable(pkg, tmp_path):
    """include/qgmap.h must be consumable from the reference's FFI language: a C99 program (examples/host_calls.c) compiles with
    -pedantic -Werror against the header, links the shared library, exercises the host-side entry points and sees a compute call
    either work (GPU box) or refuse with QGMAP_ERR_CUDA and the 'no CPU fallback' text."""
    exe = str(tmp_path / "host_calls")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "host_calls.c"), "-L" + os.path.dirname(pkg.LIB_PATH), "-lqgmap", "-lm", "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(pkg.LIB_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "host calls ok" in r.stdout and ("create refused" in r.stdout or "create ok" in r.stdout)
