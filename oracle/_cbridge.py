"""ctypes bridge to oracle/_build/libregion_oracle.so (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libregion_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "region_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def load(required: bool = False):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        try:
            build()
        except Exception:
            if required:
                raise
            return None
    lib = C.CDLL(_SO)
    i64, f32p, i64p, i32p = C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    lib.oracle_nms_sorted.restype = i64
    lib.oracle_nms_sorted.argtypes = [f32p, i64p, i64, C.c_double, i64p]
    lib.oracle_nms_sorted_batch.restype = None
    lib.oracle_nms_sorted_batch.argtypes = [f32p, i64, i64, C.c_double, i64, i64p, i64p]
    lib.oracle_roi_pool_fwd.restype = None
    lib.oracle_roi_pool_fwd.argtypes = [f32p, f32p, i64, i64, i64, i64, C.c_int, C.c_int, C.c_float, f32p, i32p]
    lib.oracle_roi_pool_bwd.restype = None
    lib.oracle_roi_pool_bwd.argtypes = [f32p, i32p, f32p, i64, i64, i64, i64, C.c_int, C.c_int, f32p]
    lib.oracle_roi_align_fwd.restype = None
    lib.oracle_roi_align_fwd.argtypes = [f32p, f32p, i64, i64, i64, i64, C.c_int, C.c_int, C.c_float, C.c_int,
                                         C.c_int, f32p]
    lib.oracle_roi_align_bwd.restype = None
    lib.oracle_roi_align_bwd.argtypes = [f32p, f32p, i64, i64, i64, i64, C.c_int, C.c_int, C.c_float, C.c_int,
                                         C.c_int, f32p]
    _lib = lib
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def nms_sorted(lib, boxes, order, thr):
    n = boxes.shape[0]
    order = np.ascontiguousarray(order, dtype=np.int64)
    keep = np.empty(n, dtype=np.int64)
    nk = lib.oracle_nms_sorted(_p(boxes, C.c_float), _p(order, C.c_int64), n, thr, _p(keep, C.c_int64))
    return keep[:nk].copy()


def nms_sorted_batch(lib, boxes, thr, max_keep):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    B, n = boxes.shape[0], boxes.shape[1]
    keep = np.full((B, max_keep), -1, dtype=np.int64)
    cnt = np.zeros(B, dtype=np.int64)
    lib.oracle_nms_sorted_batch(_p(boxes, C.c_float), B, n, thr, max_keep, _p(keep, C.c_int64), _p(cnt, C.c_int64))
    return keep, cnt


def roi_pool_fwd(lib, feat, rois5, pooled, scale):
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    rois5 = np.ascontiguousarray(rois5, dtype=np.float32).reshape(-1, 5)
    B, Cc, H, W = feat.shape
    K = rois5.shape[0]
    out = np.empty((K, Cc, pooled[0], pooled[1]), dtype=np.float32)
    arg = np.empty((K, Cc, pooled[0], pooled[1]), dtype=np.int32)
    lib.oracle_roi_pool_fwd(_p(feat, C.c_float), _p(rois5, C.c_float), K, Cc, H, W, pooled[0], pooled[1], scale,
                            _p(out, C.c_float), _p(arg, C.c_int32))
    return out, arg


def roi_pool_bwd(lib, grad_out, argmax, rois5, feat_shape):
    grad_out = np.ascontiguousarray(grad_out, dtype=np.float32)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    rois5 = np.ascontiguousarray(rois5, dtype=np.float32).reshape(-1, 5)
    B, Cc, H, W = feat_shape
    K, _, PH, PW = grad_out.shape
    gin = np.zeros(feat_shape, dtype=np.float32)
    lib.oracle_roi_pool_bwd(_p(grad_out, C.c_float), _p(argmax, C.c_int32), _p(rois5, C.c_float), K, Cc, H, W, PH,
                            PW, _p(gin, C.c_float))
    return gin


def roi_align_fwd(lib, feat, rois5, pooled, scale, sampling, aligned):
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    rois5 = np.ascontiguousarray(rois5, dtype=np.float32).reshape(-1, 5)
    B, Cc, H, W = feat.shape
    K = rois5.shape[0]
    out = np.empty((K, Cc, pooled[0], pooled[1]), dtype=np.float32)
    lib.oracle_roi_align_fwd(_p(feat, C.c_float), _p(rois5, C.c_float), K, Cc, H, W, pooled[0], pooled[1], scale,
                             sampling, int(aligned), _p(out, C.c_float))
    return out


def roi_align_bwd(lib, grad_out, rois5, feat_shape, pooled, scale, sampling, aligned):
    grad_out = np.ascontiguousarray(grad_out, dtype=np.float32)
    rois5 = np.ascontiguousarray(rois5, dtype=np.float32).reshape(-1, 5)
    B, Cc, H, W = feat_shape
    K = rois5.shape[0]
    gin = np.zeros(feat_shape, dtype=np.float32)
    lib.oracle_roi_align_bwd(_p(grad_out, C.c_float), _p(rois5, C.c_float), K, Cc, H, W, pooled[0], pooled[1], scale,
                             sampling, int(aligned), _p(gin, C.c_float))
    return gin
