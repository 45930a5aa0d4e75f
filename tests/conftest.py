import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (loads libqgmap.so; raises if it is not built)."""
    return importlib.import_module("gqmap-opticalflow_b200")


@pytest.fixture(scope="session")
def O():
    """The CPU fp64 oracle (test infrastructure)."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def T():
    from oracle import numpy_twin
    return numpy_twin


def make_problem(O, Mo, No, L, K, *, super=False, seed=0, T=0.0, lambdas=5.0, rho=0.6, small_sigma=False,
                 minu=-3.0, maxu=2.0, minv=-1.5, maxv=4.0):
    """Random textured frame pair + a random state with non-trivial correlations (shared by CPU and GPU tests)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:Mo, 0:No]
    I1 = 127 + 60 * np.sin(xx / 3.1) * np.cos(yy / 4.3) + 40 * rng.random((Mo, No))
    I2 = 127 + 60 * np.sin((xx - 1.3) / 3.1) * np.cos((yy + 0.7) / 4.3) + 40 * rng.random((Mo, No))
    I1, I2 = np.asfortranarray(I1), np.asfortranarray(I2)
    cfg = O.make_config(Mo, No, L, K, super=super, lambdas=lambdas, minu=minu, maxu=maxu, minv=minv, maxv=maxv)
    st = O.init_state(cfg, seed + 1, T=T)
    if small_sigma:
        st.sigu[:] = rng.uniform(0.05, 1.5, st.sigu.shape)
        st.sigv[:] = rng.uniform(0.05, 1.5, st.sigv.shape)
    st.pn[:] = rng.uniform(-rho, rho, st.pn.shape)
    st.rou[:] = rng.uniform(-rho, rho, st.rou.shape)
    return cfg, I1, I2, st


def options_from_cfg(cfg, its=100, T=0.0, **extra):
    o = dict(K=cfg.K, L=cfg.L, temperature=T, drate=cfg.drate, epsn=cfg.epsn, lambdad=cfg.lambdad, lambdas=cfg.lambdas,
             minu=cfg.minu, maxu=cfg.maxu, minv=cfg.minv, maxv=cfg.maxv, its=its)
    o.update(extra)
    return o


def state_dict(st):
    return dict(muu=st.muu, muv=st.muv, sigmau=st.sigu, sigmav=st.sigv, pn=st.pn, rou=st.rou, w=st.w)
