"""End-to-end probe of BASELINE configs[3] through the reference-facing call: gqmap_gpu_mixture(options, I1, I2) on one synthetic
3840x2160 pair (L=3, K=5) with options.devices = all visible GPUs (one row band per GPU, one host process).
usage: e2e_4k_devices.py [its] [ndev]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
pkg = importlib.import_module("gqmap-opticalflow_b200")
its = int(sys.argv[1]) if len(sys.argv) > 1 else 900
ndev = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
M, N = 2160, 3840
I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N)
opts = dict(K=5, L=3, its=its, temperature=0.0, drate=0.5, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv,
            seed=1, trueFlow=flow, unknownIdx=np.zeros((M, N), bool))
for devs in ([0], list(range(ndev))):
    o = dict(opts) if len(devs) == 1 else dict(opts, devices=devs)
    t0 = time.time()
    mu, sigma, alpha, AEPE, Energy, logP = pkg.gqmap_gpu_mixture(o, I1, I2)
    dt = time.time() - t0
    nl, ms = pkg.last_solve_stats()
    a = AEPE[~np.isnan(AEPE)]
    print("devices %s: %d iterations, wall %.2f s (iteration kernels %.2f s) -> %.3f Gpx-it/s end to end, %.3f Gpx-it/s in the iteration loop; AEPE %.3f -> %.3f, Energy %.6e"
          % (devs, its, dt, ms / 1e3, M * N * its / dt / 1e9, M * N * its / (ms / 1e3) / 1e9, a[0], a[-1], Energy[its - 1, 0]), flush=True)
    if len(devs) == 1:
        ref = (mu, sigma)
    else:
        print("  beliefs bit-identical to the single-GPU call:", np.array_equal(ref[0], mu) and np.array_equal(ref[1], sigma), flush=True)
