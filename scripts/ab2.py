"""A/B device-time probe of the iteration kernels (development aid): ms/iteration at a converged and at the initial state,
for kernel selections / strip heights given as env settings.  usage: ab2.py <tag> [case,case,...] [sel;sel;...]
  case = variant:M:N:L:K:burn[:g]  (g = frames rounded to integer grey levels)     sel = name=ENV1:val,ENV2:val   (library chosen by QGMAP_LIB_PATH)"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gqmap-opticalflow_b200")
tag = sys.argv[1] if len(sys.argv) > 1 else ""
cases = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] else "full:480:640:3:5:2000,full:480:640:3:5:0,full:2160:3840:3:5:300,full:388:584:1:3:1000,full:480:640:2:9:4000"
sels = sys.argv[3] if len(sys.argv) > 3 else "walk=;tile=QGMAP_ITER:tile"
frames = {}
for case in cases.split(","):
    parts = case.split(":")
    variant, (M, N, L, K, burn), grey = parts[0], map(int, parts[1:6]), len(parts) > 6 and parts[6] == "g"
    if (M, N, grey) not in frames:
        frames[(M, N, grey)] = pkg.synthetic_pair(M, N, grey_levels=grey)
    I1, I2, flow, (minu, maxu, minv, maxv) = frames[(M, N, grey)]
    opts = dict(K=K, L=L, temperature=0.2 if variant == "super" else 0.0, drate=0.75, epsn=1e-6, lambdad=1.0,
                lambdas=16.0 if variant == "super" else 5.0, minu=minu, maxu=maxu, minv=minv, maxv=maxv, its=10**6)
    for sel in sels.split(";"):
        name, _, envs = sel.partition("=")
        saved = {}
        for kv in filter(None, envs.split(",")):
            k, v = kv.split(":")
            saved[k] = os.environ.get(k)
            os.environ[k] = v
        try:
            with pkg.Solver(opts, I1, I2, variant=variant) as s:
                s.init_state(1)
                if burn:
                    s.step(burn)
                s.step(20)
                n = 200 if M < 1000 else 40
                best = 1e9
                for _ in range(3):
                    r = s.step(n)
                    best = min(best, r["ms"] / n)
                print("%-8s %-10s %-5s %4dx%-4d L=%d K=%2d burn=%4d%s: %8.4f ms/it  %7.3f Gpx-it/s  E=%.9e" % (
                    tag, name, variant, M, N, L, K, burn, " grey" if grey else "     ", best, M * N / best / 1e6, r["Energy"][-1]), flush=True)
        except Exception as e:
            print("%-8s %-10s %s FAILED: %s" % (tag, name, case, e), flush=True)
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
