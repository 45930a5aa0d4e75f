"""Run the reference's own get_map_mex / flowToColor_mex machine code (TEST INFRASTRUCTURE, see peload.c, peload_ftc.c).  Only usable
where the reference tree is present (/root/reference in the build container); elsewhere tests use the vectors committed in
tests/golden/get_map_refbin.npz and tests/golden/flow_to_color_refbin.npz."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.dirname(_HERE)
REF_BINARY = os.environ.get("QGMAP_REF_GET_MAP", "/root/reference/get_map_mex.mexw64")
REF_FTC_BINARY = os.environ.get("QGMAP_REF_FLOWTOCOLOR", "/root/reference/flowToColor_mex.mexw64")
_lib = None
_lib_ftc = None


def available():
    return os.path.exists(REF_BINARY)


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _ORACLE, "ref"], stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(_ORACLE, "_ref", "libqref.so"))
        L.qref_error.restype = C.c_char_p
        if L.qref_load(REF_BINARY.encode()) != 0:
            raise RuntimeError("cannot load the reference binary: %s" % L.qref_error().decode())
        _lib = L
    return _lib


def get_map_mex(alf, mu_u, sig_u, mu_v, sig_v):
    """map = get_map_mex(alf, mu_u, sig_u, mu_v, sig_v) computed by the reference binary's findmax (VA 0x180001c20) per pixel and
    layer -- the serial equivalent of its OpenMP loop."""
    arrs = [np.asfortranarray(a, dtype=np.float64) for a in (mu_u, sig_u, mu_v, sig_v)]
    arrs = [a.reshape(a.shape + (1,), order="F") if a.ndim == 2 else a for a in arrs]
    M, N, L = arrs[0].shape
    al = np.ascontiguousarray(np.asarray(alf, dtype=np.float64).ravel())
    assert al.size == L and 1 <= L <= 10
    out = np.zeros((M, N, 2), order="F")
    P = C.POINTER(C.c_double)
    rc = lib().qref_get_map(al.ctypes.data_as(P), *(a.ctypes.data_as(P) for a in arrs), M, N, L, out.ctypes.data_as(P))
    if rc != 0:
        nm = C.create_string_buffer(128)
        n = lib().qref_unexpected_calls(nm, 128)
        raise RuntimeError("reference binary took an error path (%d calls, last import %s)" % (n, nm.value.decode()))
    return out


def ftc_available():
    return os.path.exists(REF_FTC_BINARY)


def lib_ftc():
    global _lib_ftc
    if _lib_ftc is None:
        subprocess.check_call(["make", "-C", _ORACLE, "ref"], stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(_ORACLE, "_ref", "libqref_ftc.so"))
        L.qref_ftc_error.restype = C.c_char_p
        if L.qref_ftc_load(REF_FTC_BINARY.encode()) != 0:
            raise RuntimeError("cannot load the reference binary: %s" % L.qref_ftc_error().decode())
        _lib_ftc = L
    return _lib_ftc


def flowToColor_mex(flow):
    """[img, flo, minu, maxu, minv, maxv, idxUnknown] = flowToColor_mex(flow) computed by the reference binary's exported entry point
    `flowToColor` (legacy/flowToColor.m + legacy/computeColor.m as compiled by MATLAB Coder; call form of optical_flow.m:12-13)."""
    flow = np.asfortranarray(flow, dtype=np.float64)
    assert flow.ndim == 3 and flow.shape[2] == 2
    M, N, _ = flow.shape
    img = np.zeros((M, N, 3), dtype=np.uint8, order="F")
    flo = np.zeros((M, N, 2), order="F")
    rng = np.zeros(4)
    unk = np.zeros((M, N), dtype=np.uint8, order="F")
    V = C.c_void_p
    rc = lib_ftc().qref_flow_to_color(flow.ctypes.data_as(V), M, N, img.ctypes.data_as(V), flo.ctypes.data_as(V), rng.ctypes.data_as(V),
                                      unk.ctypes.data_as(V))
    if rc != 0:
        nm = C.create_string_buffer(128)
        n = lib_ftc().qref_ftc_unexpected_calls(nm, 128)
        raise RuntimeError("reference binary failed (rc %d; %d error-path calls, last import %s)" % (rc, n, nm.value.decode()))
    return img, flo, rng[0], rng[1], rng[2], rng[3], unk.astype(bool)
