"""ctypes wrapper + build recipe for ``oracle/postproc_c.c`` (TEST INFRASTRUCTURE)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "postproc_c.c")
LIB = os.path.join(_HERE, "_build", "liboracle_postproc.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC])
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_kth_value_topk.restype = ctypes.c_int
        _lib.oracle_greedy_nms.restype = ctypes.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def remove_borders(score, b):
    s = _f32(score)
    out = np.empty_like(s)
    lib().oracle_remove_borders(_p(s), _p(out), s.shape[0], s.shape[1], int(b))
    return out


def apply_nms(score, size):
    s = _f32(score)
    out = np.empty_like(s)
    lib().oracle_apply_nms(_p(s), _p(out), s.shape[0], s.shape[1], int(size))
    return out


def kth_value_topk(score, k):
    """-> (raster indices int32 [<=k], threshold).  IndexError like the reference if h*w < k."""
    s = _f32(score)
    idx = np.empty(max(k, 1), np.int32)
    thr = ctypes.c_float()
    n = lib().oracle_kth_value_topk(_p(s), s.shape[0], s.shape[1], int(k), _p(idx), ctypes.byref(thr))
    if n < 0:
        raise IndexError("score map has fewer than k elements")
    return idx[:n].copy(), thr.value


def greedy_nms_map(score, thr, radius):
    """threshold + nms_fast on a dense map -> surviving raster indices, score-descending."""
    s = _f32(score)
    idx = np.empty(s.size, np.int32)
    n = lib().oracle_greedy_nms(_p(s), s.shape[0], s.shape[1], ctypes.c_float(np.float32(thr)), int(radius),
                                _p(idx), idx.size)
    return idx[:n].copy()


def greedy_nms(xs, ys, scores, h, w, radius):
    """Drop-in for ``oracle.postproc.greedy_nms`` (same signature / result) when the candidates
    are distinct pixels listed in raster order, which is what the pipeline produces."""
    dense = np.full((h, w), -np.inf, np.float32)
    dense[ys, xs] = scores
    keep = greedy_nms_map(dense, -3.0e38, radius)
    lut = np.full(h * w, -1, np.int64)
    lut[np.asarray(ys, np.int64) * w + np.asarray(xs, np.int64)] = np.arange(len(xs))
    return lut[keep]
