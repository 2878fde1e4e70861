"""ctypes front-end of oracle/nms_oracle.c.  TEST INFRASTRUCTURE ONLY (see that file's header)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

_LIBS = {}


def _lib(naive=False):
    key = "naive" if naive else "fast"
    if key not in _LIBS:
        _build.build()
        name = "libnms_oracle_naive.so" if naive else "libnms_oracle.so"
        lib = ctypes.CDLL(os.path.join(_build.OUT, name))
        lib.nms_oracle_decode.restype = ctypes.c_int
        lib.nms_oracle_max_threads.restype = ctypes.c_int
        _LIBS[key] = lib
    return _LIBS[key]


def max_threads() -> int:
    return int(_lib().nms_oracle_max_threads())


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty)) if a is not None else None


def decode(proto, z, xa, sharing, weights, T, decoding_type=2, q_bit=5, clip_llr=20.0,
           want_all=True, want_c2v=False, nthreads=0, naive=False):
    """Same contract as oracle.nms_oracle.decode; returns dict(app [T,B,N*z] (or None),
    app_last [B,N*z], synd [T,B] bool, c2v [T,B,E,z] (optional))."""
    proto = np.ascontiguousarray(proto, dtype=np.int32)
    M, N = proto.shape
    xa = np.ascontiguousarray(xa, dtype=np.float32)
    B = xa.shape[0]
    E = int((proto != -1).sum())
    sh = (ctypes.c_int * 3)(*[int(s) for s in sharing])
    w = [None, None, None]
    for i in range(3):
        if sharing[i] > 0:
            w[i] = np.ascontiguousarray(np.asarray(weights[i], dtype=np.float32)[:T])
    app_all = np.empty((T, B, N * z), dtype=np.float32) if want_all else None
    app_last = np.empty((B, N * z), dtype=np.float32)
    synd = np.empty((T, B), dtype=np.uint8)
    c2v = np.empty((T, B, E, z), dtype=np.float32) if want_c2v else None
    rc = _lib(naive).nms_oracle_decode(
        _p(proto, ctypes.c_int32), M, N, int(z), sh, _p(w[0], ctypes.c_float),
        _p(w[1], ctypes.c_float), _p(w[2], ctypes.c_float), int(T), int(decoding_type), int(q_bit),
        ctypes.c_float(clip_llr), _p(xa, ctypes.c_float), B, _p(app_all, ctypes.c_float),
        _p(app_last, ctypes.c_float), _p(synd, ctypes.c_uint8), _p(c2v, ctypes.c_float), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"nms_oracle_decode failed with {rc}")
    out = {"app": app_all, "app_last": app_last, "synd": synd.astype(bool)}
    if want_c2v:
        out["c2v"] = c2v
    return out
