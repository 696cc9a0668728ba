"""ctypes access to the plain-C fp64 edge-attention oracle (oracle/gat_edge_ref.c); test-only."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libgat_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", _HERE], check=True)
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def gat_edge_fwd(XL, XR, att, bias, target, T, H, C, slope=0.2):
    lib = _load()
    XL = np.ascontiguousarray(XL, np.float32); XR = np.ascontiguousarray(XR, np.float32)
    att = np.ascontiguousarray(att, np.float32).reshape(-1)
    bias = None if bias is None else np.ascontiguousarray(bias, np.float32)
    target = np.ascontiguousarray(target, np.int64)
    E = XL.shape[0]
    out = np.zeros((T, H * C)); smax = np.zeros((T, H)); ssum = np.zeros((T, H))
    rc = lib.gat_edge_fwd_ref(_p(XL), _p(XR), ctypes.c_int(int(XR.shape[0] == 1 and T != 1)), _p(att), _p(bias), _p(target),
                              ctypes.c_int64(E), ctypes.c_int64(T), ctypes.c_int(H), ctypes.c_int(C),
                              ctypes.c_double(slope), _p(out), _p(smax), _p(ssum))
    assert rc == 0
    return out, smax, ssum


def gat_edge_bwd(XL, XR, att, target, dOut, T, H, C, slope=0.2):
    lib = _load()
    XL = np.ascontiguousarray(XL, np.float32); XR = np.ascontiguousarray(XR, np.float32)
    att = np.ascontiguousarray(att, np.float32).reshape(-1)
    target = np.ascontiguousarray(target, np.int64); dOut = np.ascontiguousarray(dOut, np.float32)
    E = XL.shape[0]
    dXL = np.zeros((E, H * C)); dXR = np.zeros((T, H * C)); datt = np.zeros(H * C); dbias = np.zeros(H * C)
    rc = lib.gat_edge_bwd_ref(_p(XL), _p(XR), ctypes.c_int(int(XR.shape[0] == 1 and T != 1)), _p(att), _p(target), _p(dOut),
                              ctypes.c_int64(E), ctypes.c_int64(T), ctypes.c_int(H), ctypes.c_int(C),
                              ctypes.c_double(slope), _p(dXL), _p(dXR), _p(datt), _p(dbias))
    assert rc == 0
    if XR.shape[0] == 1 and T != 1:
        dXR = dXR.sum(axis=0, keepdims=True)
    return dXL, dXR, datt, dbias
