"""ctypes binding of ``cmatrices_oracle.c`` (the C restatement of pyradiomics'
``cmatrices.c``).  TEST INFRASTRUCTURE ONLY — imported by tests/, smoke() and
bench.py's CPU-baseline legs, never by the product package."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_cmatrices.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "cmatrices_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle_cmatrices.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def glcm(levels, Ng, uni, symmetrical=True):
    lev = _i32(levels)
    ang = _i32(uni).reshape(-1, 2)
    out = np.zeros((Ng, Ng, len(ang)), dtype=np.int64)
    lib().orc_glcm(_p(lev), lev.shape[0], lev.shape[1], Ng, _p(ang), len(ang), int(bool(symmetrical)), _p(out))
    return out


def glrlm(levels, Ng, uni):
    lev = _i32(levels)
    ang = _i32(uni).reshape(-1, 2)
    Nr = max(lev.shape)
    out = np.zeros((Ng, Nr, len(ang)), dtype=np.int64)
    lib().orc_glrlm(_p(lev), lev.shape[0], lev.shape[1], Ng, Nr, _p(ang), len(ang), _p(out))
    return out


def glszm(levels, Ng, bi):
    lev = _i32(levels)
    ang = _i32(bi).reshape(-1, 2)
    Ns = max(int((lev > 0).sum()), 1)
    out = np.zeros((Ng, Ns), dtype=np.int64)
    rc = lib().orc_glszm(_p(lev), lev.shape[0], lev.shape[1], Ng, Ns, _p(ang), len(ang), _p(out))
    if rc:
        raise MemoryError
    return out


def gldm(levels, Ng, bi, alpha=0):
    lev = _i32(levels)
    ang = _i32(bi).reshape(-1, 2)
    out = np.zeros((Ng, len(ang) + 1), dtype=np.int64)
    lib().orc_gldm(_p(lev), lev.shape[0], lev.shape[1], Ng, _p(ang), len(ang), int(alpha), _p(out))
    return out


def ngtdm(levels, Ng, bi):
    lev = _i32(levels)
    ang = _i32(bi).reshape(-1, 2)
    n = np.zeros(Ng, dtype=np.int64)
    s = np.zeros(Ng, dtype=np.float64)
    lib().orc_ngtdm(_p(lev), lev.shape[0], lev.shape[1], Ng, _p(ang), len(ang), _p(n), _p(s))
    return n, s


def bin_image(image, roi, binWidth):
    img = np.ascontiguousarray(image, dtype=np.float64)
    r = np.ascontiguousarray(roi, dtype=np.uint8)
    lev = np.zeros(img.shape, dtype=np.int32)
    f = lib().orc_bin_image
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    Ng = f(_p(img), _p(r), img.shape[0], img.shape[1], float(binWidth), _p(lev))
    return lev, Ng
