"""gqmap-opticalflow_b200: B200-native (sm_100a CUDA) QGMAP optical-flow inference behind the reference's
MATLAB-level interface.  Importing the package loads libqgmap.so and fails loudly if it is missing.

The directory name carries a hyphen (it mirrors the reference repo name); import it with
    importlib.import_module("gqmap-opticalflow_b200")
"""
from . import _lib                                     # noqa: F401  (raises if libqgmap.so is absent)
from ._lib import QgmapError, QgmapConfig, LIB_PATH    # noqa: F401
from .host import (gqmap_gpu_mixture, gqmap_gpuSuper_mix_entropy, get_map_mex, flowToColor_mex,   # noqa: F401
                   GaussHermite_2, projsplx, imwrite, Solver, make_config, last_solve_stats, batch_step, fp32_peak, BandGroup)
from .frames import readFlowFile, writeFlowFile, save_results, rgb2gray, synthetic_pair, middlebury_shapes                     # noqa: F401
from . import dist                                    # noqa: F401,E402
from . import ctf                                     # noqa: F401,E402
from .ctf import optical_flow_ctf, gqmap_ctf          # noqa: F401,E402
