"""Golden vectors for get_map_mex produced by the REFERENCE'S OWN BINARY (get_map_mex.mexw64, run through oracle/refbin/peload.c).
Run in the build container (needs /root/reference):  python tests/golden/make_refbin_golden.py  ->  tests/golden/get_map_refbin.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refbin import refbin


def cases():
    rng = np.random.default_rng(20181003)                 # the binary's build date
    M, N = 12, 10
    for L in range(1, 11):
        for kind in ("tight", "medium", "wide", "ties"):
            al = rng.random(L) + 0.05
            al /= al.sum()
            spread = dict(tight=0.3, medium=2.0, wide=8.0, ties=1.0)[kind]
            mu_u, mu_v = rng.normal(0, spread, (M, N, L)), rng.normal(0, spread, (M, N, L))
            sig_u, sig_v = rng.uniform(0.05, 3.0, (M, N, L)), rng.uniform(0.05, 3.0, (M, N, L))
            if kind == "ties":                                   # repeated means, equal weights: the strict-< / first-index rules
                mu_u[:, :, L // 2:] = mu_u[:, :, :1]
                sig_v[:, :, :] = sig_v[:, :, :1]
                al[:] = 1.0 / L
            yield "L%d_%s" % (L, kind), al, mu_u, sig_u, mu_v, sig_v


if __name__ == "__main__":
    out = {}
    for name, al, mu_u, sig_u, mu_v, sig_v in cases():
        m = refbin.get_map_mex(al, mu_u, sig_u, mu_v, sig_v)
        out.update({name + "_alpha": al, name + "_mu_u": mu_u, name + "_sig_u": sig_u, name + "_mu_v": mu_v, name + "_sig_v": sig_v,
                    name + "_map": m})
    np.savez_compressed(os.path.join(HERE, "get_map_refbin.npz"), **out)
    print("wrote", len(out) // 6, "cases,", os.path.getsize(os.path.join(HERE, "get_map_refbin.npz")) // 1024, "KiB")
