"""ctypes wrapper around oracle/lsap.c (CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblsap_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "lsap.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.lsap_f32.restype = ctypes.c_int
        _lib.lsap_f32_batch.restype = ctypes.c_int
    return _lib


def lsap(cost):
    """cost [nr, nc] float32 -> (rows int64 ascending, cols int64), like scipy."""
    cost = np.ascontiguousarray(cost, dtype=np.float32)
    nr, nc = cost.shape
    k = min(nr, nc)
    rows = np.zeros(max(k, 1), np.int64)
    cols = np.zeros(max(k, 1), np.int64)
    rc = _load().lsap_f32(cost.ctypes.data_as(ctypes.c_void_p), nr, nc,
                          rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p))
    if rc == -1:
        raise ValueError("matrix contains invalid numeric entries")
    if rc == -2:
        raise ValueError("cost matrix is infeasible")
    return rows[:k], cols[:k]


def lsap_batch(cost, n):
    """cost [B, Q, ld] float32, n [B] int32 -> rows, cols [B, Q] int64 (first n[b] valid)."""
    cost = np.ascontiguousarray(cost, dtype=np.float32)
    B, Q, ld = cost.shape
    n = np.ascontiguousarray(n, dtype=np.int32)
    rows = np.zeros((B, Q), np.int64)
    cols = np.zeros((B, Q), np.int64)
    rc = _load().lsap_f32_batch(cost.ctypes.data_as(ctypes.c_void_p), B, Q,
                                n.ctypes.data_as(ctypes.c_void_p), ld,
                                rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise ValueError("lsap failed rc=%d" % rc)
    return rows, cols
