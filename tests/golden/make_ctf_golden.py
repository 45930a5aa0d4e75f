"""Golden vector produced by EXECUTING THE REFERENCE'S legacy/gqmap_ctf.m (the per-level solver of the coarse-to-fine driver,
legacy/optical_flow_ctf.m:21-36), unmodified and read from /root/reference, under oracle/mlab/minimat.py.
The interpreter is frozen; what gqmap_ctf.m needs beyond it comes in as externals of THIS script:
  interp2(V,k,'cubic')  MATLAB toolbox code (not shipped): restated by oracle.interp2_cubic_refine -- getVV + the node_pot weights,
                        i.e. what the reference itself hand-copies from interp2.m; so the 64x upsampled frame is OUR restatement
                        (unpinned third party), while everything gqmap_ctf.m does with it is executed source;
  imshow                a no-op;     flowToColor -> legacy/flowToColor.m is found on the interpreter's path and executed.
Run in the build container:  python tests/golden/make_ctf_golden.py  ->  tests/golden/refsrc_ctf.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference"
Mo, No, K, ITS, SEED = 24, 32, 3, 3, 31


def inputs():
    rng = np.random.default_rng(SEED)
    yy, xx = np.mgrid[0:Mo, 0:No]
    I1 = np.round(127 + 60 * np.sin(xx / 2.1) * np.cos(yy / 2.3) + 40 * rng.random((Mo, No)))
    I2 = np.round(127 + 60 * np.sin((xx - 0.8) / 2.1) * np.cos((yy + 0.5) / 2.3) + 40 * rng.random((Mo, No)))
    grdt = np.stack([1.5 * np.sin(yy / 7.0) + 0.3, 0.8 * np.cos(xx / 9.0)], axis=2)
    return np.asfortranarray(I1), np.asfortranarray(I2), np.asfortranarray(grdt)


def main():
    from oracle import oracle as O
    from oracle.mlab.minimat import Interp
    I1, I2, grdt = inputs()
    rng = np.random.default_rng(SEED + 1000)
    draws, snaps = [], {}

    def rand(shape):
        a = rng.random(int(np.prod(shape))).reshape(shape, order="F")
        draws.append(np.array(a))
        return a

    def probe(ws):
        it = int(ws["it"])
        snaps[it] = {f: np.array(np.asarray(ws[f], dtype=np.float64), order="F") for f in ("muu", "muv", "sigmau", "sigmav", "pn", "rou")}
        snaps[it].update(ptdmu=float(ws["ptdmu"]), ptdsigma=float(ws["ptdsigma"]), aepe=float(ws["aepe"]))

    def interp2(nargout, V, k, method):
        assert str(method) == "cubic"
        return (O.interp2_cubic_refine(np.asarray(V, dtype=np.float64), int(k)),)
    interp = Interp([os.path.join(REF, "legacy"), REF], rand=rand, on_fprintf=probe,
                    externals={"interp2": interp2, "imshow": lambda n, *a: ()})
    opts = dict(its=float(ITS), K=float(K), epsn=1e-6, lambdad=1.0, lambdas=5.0)
    mu, sigma, rou, AEPE, Energy = interp.call("gqmap_ctf", opts, I1, I2, grdt, nargout=5)
    out = dict(I1=I1, I2=I2, GRDT=grdt, mu=mu, sigma=sigma, rou=rou, AEPE=np.ravel(AEPE), Energy=np.ravel(Energy),
               pn=np.asarray(interp.last_workspace["pn"]), meta=np.array([Mo, No, K, ITS]))
    for i, d in enumerate(draws):
        out["draw%d" % i] = d
    for it, d in snaps.items():
        for k, v in d.items():
            out["p%d_%s" % (it, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "refsrc_ctf.npz"), **out)
    print("gqmap_ctf.m executed: %d iterations, Energy %s, AEPE %s, %d rand draws" % (ITS, np.ravel(Energy), np.ravel(AEPE)[:ITS], len(draws)))


if __name__ == "__main__":
    main()
