"""Row bands over several GPUs: multi-process (one process per GPU; peer-memory publish kernel and NCCL transports) and the
single-process band group with one band per GPU.  Needs >= 2 GPUs: run with `gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_nccl_bands_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (covered on one GPU by test_gpu_bands.py through the same iteration structure)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29621", os.path.join("tests", "_nccl_band_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("bit-identical") == 5


@pytest.mark.gpu
def test_group_one_band_per_gpu_is_bit_identical():
    """qgmap_group with every band on its own GPU uses the peer-memory publish kernel (qgmap_p2p.cu): same bits as one domain."""
    import importlib
    import numpy as np
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("gqmap-opticalflow_b200")
    from oracle import oracle as O
    from conftest import make_problem, options_from_cfg, state_dict
    nb = min(ndev, 4)
    for variant, (Mo, No), L, K, T in (("full", (67, 90), 2, 5, 0.0), ("super", (128, 160), 3, 3, 0.2)):
        cfg, I1, I2, st = make_problem(O, Mo, No, L, K, super=variant == "super", seed=29, T=T, small_sigma=True)
        opts = options_from_cfg(cfg, T=T, alpha_scale=1e-5)
        with pkg.Solver(opts, I1, I2, variant=variant) as s1:
            s1.set_state(state_dict(st), T=T, it=495)
            r1 = s1.step(14)
            a = s1.get_state()
        with pkg.BandGroup(opts, I1, I2, nb, devices=list(range(nb)), variant=variant) as g:
            g.set_state(state_dict(st), T=T, it=495)
            ra = g.step(6)
            rb = g.step(8)
            b = g.get_state()
        for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou"):
            assert np.array_equal(a[f], b[f]), (variant, f)
        E = np.concatenate([ra["Energy"], rb["Energy"]])
        assert np.abs(E / r1["Energy"] - 1).max() < 1e-12 and np.abs(a["alpha"] - b["alpha"]).max() < 1e-14


@pytest.mark.gpu
def test_solve_with_options_devices_on_distinct_gpus():
    """The reference-facing call with options.devices = [0,1,...]: one band per GPU, peer-memory exchange kernel underneath."""
    import importlib
    import numpy as np
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("gqmap-opticalflow_b200")
    Mo, No = 96, 120
    I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(Mo, No)
    opts = dict(K=5, L=3, its=25, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv,
                maxv=maxv, seed=5, log_every=10, trueFlow=flow, unknownIdx=np.zeros((Mo, No), bool), alpha_start=2, alpha_scale=1e-5)
    a = pkg.gqmap_gpu_mixture(opts, I1, I2)
    b = pkg.gqmap_gpu_mixture(dict(opts, devices=list(range(min(ndev, 4)))), I1, I2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.abs(a[2] - b[2]).max() < 1e-14
    for k in (3, 4, 5):
        m = ~np.isnan(a[k])
        assert np.array_equal(np.isnan(a[k]), np.isnan(b[k])) and np.abs(b[k][m] / a[k][m] - 1).max() < 1e-11
