"""ctypes binding of libqgmap.so (include/qgmap.h).  No torch types cross this boundary.

The library is the product: if it is missing or fails to load, importing this module raises -- there is no
CPU/NumPy fallback anywhere in the package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# QGMAP_LIB_PATH: development aid for A/B timing of two builds of the same library (never a fallback: the file must exist)
LIB_PATH = os.environ.get("QGMAP_LIB_PATH") or os.path.join(_HERE, "libqgmap.so")

QGMAP_LMAX = 10
QGMAP_P2P_BLOB_BYTES = 512
QGMAP_KMAX = 32
VARIANT_FULL, VARIANT_SUPER = 0, 1
ALPHA_SOFTMAX, ALPHA_PROJSPLX = 0, 1


class QgmapError(RuntimeError):
    def __init__(self, status, text):
        super().__init__("libqgmap status %d: %s" % (status, text))
        self.status = status


class QgmapConfig(C.Structure):
    """Mirror of qgmap_config (include/qgmap.h)."""
    _fields_ = [
        ("struct_size", C.c_int32), ("variant", C.c_int32), ("L", C.c_int32), ("K", C.c_int32),
        ("lambdad", C.c_double), ("lambdas", C.c_double), ("epsn", C.c_double),
        ("temperature", C.c_double), ("drate", C.c_double),
        ("minu", C.c_double), ("maxu", C.c_double), ("minv", C.c_double), ("maxv", C.c_double),
        ("sigma_min", C.c_double), ("sigma_max", C.c_double), ("corr_tor", C.c_double),
        ("step0", C.c_double), ("step_tau", C.c_double), ("alpha_scale", C.c_double),
        ("T_floor", C.c_double), ("tor", C.c_double), ("sigma_step_scale", C.c_double),
        ("alpha_start", C.c_int32), ("alpha_mode", C.c_int32), ("anneal_every", C.c_int32),
        ("device", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32), ("log_every", C.c_int32), ("strip_rows", C.c_int32),
    ]


_DP = C.POINTER(C.c_double)
_U8P = C.POINTER(C.c_uint8)
_IP = C.POINTER(C.c_int)

# name -> (restype, argtypes): every symbol include/qgmap.h declares
SIGNATURES = {
    "qgmap_config_defaults": (C.c_int, [C.POINTER(QgmapConfig), C.c_int]),
    "qgmap_create": (C.c_int, [C.POINTER(QgmapConfig), _DP, _DP, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "qgmap_destroy": (C.c_int, [C.c_void_p]),
    "qgmap_dims": (C.c_int, [C.c_void_p, _IP, _IP, _IP]),
    "qgmap_set_state": (C.c_int, [C.c_void_p] + [_DP] * 8 + [C.c_double, C.c_int]),
    "qgmap_get_state": (C.c_int, [C.c_void_p] + [_DP] * 8 + [_DP, _IP]),
    "qgmap_set_state_f32": (C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [_DP] * 2 + [C.c_double, C.c_int]),
    "qgmap_get_state_f32": (C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [_DP] * 2 + [_DP, _IP]),
    "qgmap_init_state": (C.c_int, [C.c_void_p, C.c_uint64]),
    "qgmap_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _DP, _DP, _DP, _IP, _IP]),
    "qgmap_step_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "qgmap_step_end": (C.c_int, [C.c_void_p, _DP, _DP, _DP, _IP, _IP]),
    "qgmap_batch_step": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]),
    "qgmap_fp32_peak": (C.c_int, [C.c_int, _DP]),
    "qgmap_last_step_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "qgmap_last_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong)]),
    "qgmap_get_map": (C.c_int, [C.c_void_p, _DP]),
    "qgmap_logp": (C.c_int, [C.c_void_p, _DP, _DP]),
    "qgmap_aepe": (C.c_int, [C.c_void_p, _DP, _DP, _U8P, _DP]),
    "qgmap_image_dims": (C.c_int, [C.c_void_p, _IP, _IP]),
    "qgmap_set_truth": (C.c_int, [C.c_void_p, _DP, _U8P]),
    "qgmap_monitor_partial": (C.c_int, [C.c_void_p, _DP, _DP]),
    "qgmap_solve": (C.c_int, [C.POINTER(QgmapConfig), _DP, _DP, C.c_int, C.c_int, C.c_int, C.POINTER(_DP), C.c_uint64,
                              _DP, _U8P, _DP, _DP, _DP, _DP, _DP, _DP, _IP]),
    "qgmap_group_solve": (C.c_int, [C.POINTER(QgmapConfig), _DP, _DP, C.c_int, C.c_int, C.c_int, C.c_int, _IP, C.POINTER(_DP), C.c_uint64,
                                    _DP, _U8P, _DP, _DP, _DP, _DP, _DP, _DP, _IP]),
    "qgmap_last_solve_stats": (C.c_int, [C.POINTER(C.c_longlong), C.POINTER(C.c_float)]),
    "qgmap_find_map": (C.c_int, [_DP] * 5 + [C.c_int, C.c_int, C.c_int, _DP, C.c_int]),
    "qgmap_flow_to_color": (C.c_int, [_DP, C.c_int, C.c_int, C.c_double, _U8P, _DP, _DP, _U8P]),
    "qgmap_write_png": (C.c_int, [C.c_char_p, _U8P, C.c_int, C.c_int]),
    "qgmap_solve_set_dump_dir": (C.c_int, [C.c_char_p]),
    "qgmap_read_flo": (C.c_int, [C.c_char_p, _IP, _IP, _DP]),
    "qgmap_write_flo": (C.c_int, [C.c_char_p, _DP, C.c_int, C.c_int]),
    "qgmap_gauss_hermite": (C.c_int, [C.c_int, _DP, _DP]),
    "qgmap_projsplx": (C.c_int, [_DP, C.c_int, _DP]),
    "qgmap_debug_gradients": (C.c_int, [C.c_void_p] + [_DP] * 8),
    "qgmap_band_unique_id": (C.c_int, [C.c_void_p]),
    "qgmap_band_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "qgmap_band_p2p_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "qgmap_band_p2p_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "qgmap_group_create": (C.c_int, [C.POINTER(QgmapConfig), _DP, _DP, C.c_int, C.c_int, C.c_int, _IP, C.POINTER(C.c_void_p)]),
    "qgmap_group_destroy": (C.c_int, [C.c_void_p]),
    "qgmap_group_dims": (C.c_int, [C.c_void_p, _IP, _IP, _IP, _IP]),
    "qgmap_group_set_state": (C.c_int, [C.c_void_p] + [_DP] * 8 + [C.c_double, C.c_int]),
    "qgmap_group_init_state": (C.c_int, [C.c_void_p, C.c_uint64]),
    "qgmap_group_get_state": (C.c_int, [C.c_void_p] + [_DP] * 8 + [_DP, _IP]),
    "qgmap_group_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _DP, _DP, _DP, _IP, _IP]),
    "qgmap_group_last_step_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "qgmap_last_error": (C.c_char_p, [C.c_void_p]),
    "qgmap_status_string": (C.c_char_p, [C.c_int]),
    "qgmap_version": (C.c_int, []),
}

if not os.path.exists(LIB_PATH):
    raise ImportError("libqgmap.so not built (%s missing): run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "or `make -C gqmap-opticalflow_b200/csrc`; there is no CPU fallback" % LIB_PATH)
lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here == the .so does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args


def check(status, handle=None):
    if status != 0:
        msg = lib.qgmap_last_error(handle) or b""
        raise QgmapError(status, "%s: %s" % (lib.qgmap_status_string(status).decode(), msg.decode(errors="replace")))


def dptr(a):
    return None if a is None else a.ctypes.data_as(_DP)


def u8ptr(a):
    return None if a is None else a.ctypes.data_as(_U8P)


def f64(a):
    """MATLAB layout at the boundary: fp64, column-major."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))
