"""Bitwise comparison of the beliefs after n iterations between builds of libqgmap (development aid).
usage: bitcmp.py lib1.so lib2.so ...   (each build runs in its own process; prints a hash of every state field per build)"""
import hashlib, importlib, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    pkg = importlib.import_module("gqmap-opticalflow_b200")
    import numpy as np
    for (M, N, L, K) in ((96, 128, 3, 5), (70, 100, 2, 9)):
        I1, I2, flow, (minu, maxu, minv, maxv) = pkg.synthetic_pair(M, N, grey_levels=True)
        opts = dict(K=K, L=L, temperature=0.0, drate=0.75, epsn=1e-6, lambdad=1.0, lambdas=5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, its=10**6)
        with pkg.Solver(opts, I1, I2, variant="full") as s:
            s.init_state(1)
            for n in (1, 1, 8):
                s.step(n)
                st = s.get_state()
                print(M, N, L, K, "it", st["it"], " ".join("%s=%s" % (k, hashlib.md5(np.ascontiguousarray(st[k]).tobytes()).hexdigest()[:8])
                                                          for k in ("muu", "sigmau", "pn", "rou")), flush=True)
else:
    for lib in sys.argv[1:]:
        print("==", lib, flush=True)
        env = dict(os.environ, QGMAP_LIB_PATH=lib)
        subprocess.run([sys.executable, __file__, "--child"], env=env)
