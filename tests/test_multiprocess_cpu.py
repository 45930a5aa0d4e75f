"""world_size-2 tests on CPU (gloo): the N>1 host logic and the row-band protocol, plus the bench's rank handling."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)


def test_band_protocol_world2_gloo():
    r = _torchrun([os.path.join("tests", "_dist_worker.py")], 29611)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_bench_reference_arm_under_torchrun():
    """Under torchrun only rank 0 runs and prints the reference arm; the other rank exits 0 without work."""
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--band-rows", "135", "--band-cols", "240"], 29612)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pixel-iter/s" and d["value"] > 0 and d["n_gpus"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"
    sys.path.insert(0, ROOT)
    import bench                                     # both arms must print the same config.workload (the driver compares them)
    assert d["config"]["workload"] == bench.WORKLOAD and d["scaling"] == "strong"
