"""Tensor-level wrappers over the C ABI (include/frr.h).

PyTorch is plumbing here: it owns device memory and streams; every computation is a kernel of
libfrr.so.  All functions require CUDA fp32 tensors and raise otherwise (no CPU fallback).
Batched layouts: a leading ``B`` dimension = images; ragged results come back as
fixed-capacity tensors + ``int32 count[B]`` on the device (no host sync).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_MIN_SIZE = float(np.float32(1.0 / 1000.0))  # models/model.py:39, compared in fp32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _req(t: torch.Tensor, name: str, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor: the region stage has no CPU path")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


def anchor_base_table(base_size: int = 16) -> np.ndarray:
    """anchor.py:15-32 (host, 9x4 fp32)."""
    out = np.empty((9, 4), dtype=np.float32)
    _lib.check(_lib.load().frr_anchor_base_host(out.ctypes.data, int(base_size)), "frr_anchor_base_host")
    return out


def _table_arg(table):
    if table is None:
        return None, 0, 9
    t = np.ascontiguousarray(table, dtype=np.float32).reshape(-1, 4)
    return t, t.ctypes.data, t.shape[0]


def anchors(image_hw, device, stride: int = 16, table=None) -> torch.Tensor:
    """anchor.py:34-55 on the device: [fh*fw*A, 4] fp32, normalised."""
    lib = _lib.load()
    keep, tptr, A = _table_arg(table)
    H, W = int(image_hw[0]), int(image_hw[1])
    n = (H // stride) * (W // stride) * A
    with torch.cuda.device(device):
        out = torch.empty((n, 4), dtype=torch.float32, device=device)
        _lib.check(lib.frr_anchors(out.data_ptr(), H, W, stride, tptr, A, _stream()), "frr_anchors")
    return out


def rpn_decode(reg, cls, image_hw=None, anchors=None, stride: int = 16, table=None, min_size: float = _MIN_SIZE):
    """Fused A2+P1+P2+P3.  reg [B,N,4]; cls [B,N,2] logits or [B,N] scores.
    Returns boxes [B,N,4], scores [B,N], valid uint8 [B,N]."""
    lib = _lib.load()
    reg = _req(reg, "reg")
    cls = _req(cls, "cls")
    if reg.dim() != 3 or reg.shape[-1] != 4:
        raise ValueError("reg must be [B,N,4]")
    B, N = reg.shape[0], reg.shape[1]
    logits = cls.dim() == 3
    if tuple(cls.shape) != ((B, N, 2) if logits else (B, N)):
        raise ValueError("cls must be [B,N,2] logits or [B,N] scores")
    keep, tptr, A = _table_arg(table)
    if anchors is not None:
        anchors = _req(anchors, "anchors")
        if tuple(anchors.shape) != (N, 4):
            raise ValueError("anchors must be [N,4]")
        H = W = 0
    else:
        if image_hw is None:
            raise ValueError("image_hw is required when anchors are generated in-kernel")
        H, W = int(image_hw[0]), int(image_hw[1])
    with torch.cuda.device(reg.device):
        boxes = torch.empty((B, N, 4), dtype=torch.float32, device=reg.device)
        scores = torch.empty((B, N), dtype=torch.float32, device=reg.device)
        valid = torch.empty((B, N), dtype=torch.uint8, device=reg.device)
        _lib.check(lib.frr_rpn_decode(reg.data_ptr(), cls.data_ptr(), int(logits), _ptr(anchors), tptr, A, H, W, stride,
                                      float(min_size), boxes.data_ptr(), scores.data_ptr(), valid.data_ptr(), B, N,
                                      _stream()), "frr_rpn_decode")
    return boxes, scores, valid


def topk_desc(scores, k: int, valid=None, boxes=None, want_cidx: bool = False):
    """P4.  scores [B,N] -> dict(scores [B,k], idx int32 [B,k], cidx, boxes [B,k,4], count int32 [B])."""
    lib = _lib.load()
    scores = _req(scores, "scores")
    if scores.dim() != 2:
        raise ValueError("scores must be [B,N]")
    B, N = scores.shape
    if valid is not None:
        valid = _req(valid, "valid", torch.uint8)
    if boxes is not None:
        boxes = _req(boxes, "boxes")
    dev = scores.device
    with torch.cuda.device(dev):
        o_s = torch.empty((B, k), dtype=torch.float32, device=dev)
        o_i = torch.empty((B, k), dtype=torch.int32, device=dev)
        o_c = torch.empty((B, k), dtype=torch.int32, device=dev) if want_cidx else None
        o_b = torch.empty((B, k, 4), dtype=torch.float32, device=dev) if boxes is not None else None
        cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
        _lib.check(lib.frr_topk_desc(scores.data_ptr(), _ptr(valid), _ptr(boxes), B, N, int(k), o_s.data_ptr(),
                                     o_i.data_ptr(), _ptr(o_c), _ptr(o_b), cnt.data_ptr(), _stream()), "frr_topk_desc")
    return dict(scores=o_s, idx=o_i, cidx=o_c, boxes=o_b, count=cnt)


def nms_sorted(boxes, iou_threshold: float, max_keep: int | None = None, counts=None, gather: bool = True,
               cluster_size: int = 0, threads: int = 0, dbg=None):
    """N1 on score-sorted boxes [B,n,4].  Returns keep int32 [B,max_keep] (-1 padded), count int32 [B],
    rois [B,max_keep,4] (zero padded) or None."""
    lib = _lib.load()
    boxes = _req(boxes, "boxes")
    if boxes.dim() != 3 or boxes.shape[-1] != 4:
        raise ValueError("boxes must be [B,n,4]")
    B, n = boxes.shape[0], boxes.shape[1]
    mk = n if max_keep is None else int(max_keep)
    if counts is not None:
        counts = _req(counts, "counts", torch.int32)
    dev = boxes.device
    with torch.cuda.device(dev):
        keep = torch.empty((B, mk), dtype=torch.int32, device=dev)
        cnt = torch.empty((B,), dtype=torch.int32, device=dev)
        rois = torch.empty((B, mk, 4), dtype=torch.float32, device=dev) if gather else None
        _lib.check(lib.frr_nms_sorted_tuned(boxes.data_ptr(), _ptr(counts), B, n, float(iou_threshold), mk,
                                            keep.data_ptr(), cnt.data_ptr(), _ptr(rois), int(cluster_size), int(threads),
                                            _ptr(dbg), _stream()), "frr_nms_sorted")
    return keep, cnt, rois
