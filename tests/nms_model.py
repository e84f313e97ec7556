"""CPU model (numpy, fp32 step by step) of the algorithm of the bucketed NMS kernel `nms_bucket_kernel`
(faster_rcnn_pytorch_b200/csrc/nms_bucket.cu): large chunks of candidates, every pair test restricted to the
admissible (area class, x bin) keys, screen before the exact test, predecessor fix-point inside a chunk.

It exists to check ON THE CPU that the kernel's pruning rules are exact necessary conditions -- i.e. that the
keep list it produces equals the oracle's (= torchvision's CPU kernel) on the adversarial sets of nms_cases.py --
with the constants computed exactly as nms.cu::make_thr computes them.  Test infrastructure only."""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
K_STRIPS, K_XBINS = 88, 8
K_KEYS = K_STRIPS * K_XBINS


def make_thr(thr: float):
    """nms.cu::make_thr"""
    f = f32(thr)
    if math.isnan(thr):
        up = f32(np.nan)
    else:
        if not (float(f) > thr):
            f = np.nextafter(f, f32(np.inf))
        up = f32(f)
    fast = (thr >= 1.0e-6) and math.isfinite(thr) and bool(up < f32(1.0e30))
    u = float(up)
    c2 = f32(u / (1.0 + u) * (1.0 - 1.9073486328125e-06)) if fast else f32(0)
    tl = thr * (1.0 - 7.62939453125e-06)
    fx = f32(max(1.0 - tl, (1.0 - tl) / (2.0 * tl)) * (1.0 + 1.0e-6)) if fast else f32(0)
    alo = f32(tl * (1.0 - 1.0e-6)) if fast else f32(0)
    ahi = f32(1.0 / tl * (1.0 + 1.0e-6)) if fast else f32(0)
    return dict(up=up, c2=c2, fx=fx, alo=alo, ahi=ahi, fast=fast)


def box_area(b):
    return ((b[..., 2] - b[..., 0]).astype(f32) * (b[..., 3] - b[..., 1]).astype(f32)).astype(f32)


def screen_area(b, c2):
    a = box_area(b)
    with np.errstate(invalid="ignore"):
        ok = (b[..., 2] >= b[..., 0]) & (b[..., 3] >= b[..., 1]) & (a <= f32(3.0e38)) & (a >= f32(1.0e-30))
    return np.where(ok, (c2 * a).astype(f32), f32(np.nan)).astype(f32)


def strip_of_area(a):
    bits = np.asarray(a, f32).view(np.int32)
    return np.minimum(K_STRIPS - 1, np.maximum(0, (bits >> 21) - ((127 - 22) << 2)))


def xbin_of(cx):
    with np.errstate(invalid="ignore"):
        v = np.trunc((np.asarray(cx, f32) * f32(K_XBINS)).astype(f32))
    v = np.where(np.isnan(v), 0, v)
    return np.minimum(K_XBINS - 1, np.maximum(0, v.astype(np.int64)))


def keys_of(b, sa):
    cls = strip_of_area(box_area(b))
    xb = xbin_of((f32(0.5) * (b[..., 0] + b[..., 2]).astype(f32)).astype(f32))
    always = np.isnan(sa)
    return cls, xb, always


def admissible_ranges(b, sa, t):
    """Per candidate: class range [c_lo, c_hi], bin range [x_lo, x_hi]; `everything` for NaN screening areas."""
    a = box_area(b)
    with np.errstate(invalid="ignore", over="ignore"):
        c_lo = strip_of_area((a * t["alo"]).astype(f32))
        c_hi = strip_of_area((a * t["ahi"]).astype(f32))
        cx = (f32(0.5) * (b[..., 0] + b[..., 2]).astype(f32)).astype(f32)
        rx = (((t["fx"] * (b[..., 2] - b[..., 0]).astype(f32)).astype(f32) * f32(1.0001)).astype(f32) + f32(2.0e-6)).astype(f32)
        x_lo = xbin_of((cx - rx).astype(f32))
        x_hi = xbin_of((cx + rx).astype(f32))
    return c_lo, c_hi, x_lo, x_hi, np.isnan(sa)


def _pair_terms(a, b):
    """a [m,4] (earlier boxes), b [n,4]: fp32 w, h (unclamped difference), as the kernels compute them."""
    w = (np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0])).astype(f32)
    h = (np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1])).astype(f32)
    return w, h


def suppress_screen(a, sa, b, sb):
    """suppress_screen<true>: False only when the pair is certainly not suppressed.  [m,n]"""
    w, hd = _pair_terms(a, b)
    with np.errstate(invalid="ignore"):
        h = np.where(np.isnan(hd), f32(0), np.clip(hd, f32(0), f32(1))).astype(f32)
        return ~((w * h).astype(f32) < (sa[:, None] + sb[None, :]).astype(f32))


def suppress_exact(a, b, up):
    w, h = _pair_terms(a, b)
    w = np.maximum(f32(0), w); h = np.maximum(f32(0), h)
    inter = (w * h).astype(f32)
    with np.errstate(invalid="ignore", divide="ignore"):
        uni = ((box_area(a)[:, None] + box_area(b)[None, :]).astype(f32) - inter).astype(f32)
        ovr = (inter / uni).astype(f32)
        return ovr >= up


def _admissible(cls_l, xb_l, always_l, rng):
    """[len(list), len(cands)] mask: list member lies in the candidate's admissible keys."""
    c_lo, c_hi, x_lo, x_hi, every = rng
    m = (cls_l[:, None] >= c_lo[None, :]) & (cls_l[:, None] <= c_hi[None, :]) & \
        (xb_l[:, None] >= x_lo[None, :]) & (xb_l[:, None] <= x_hi[None, :])
    return m | always_l[:, None] | every[None, :]


def nms_bucketed(boxes: np.ndarray, thr: float, max_keep: int, chunk_sizes=(2048,), stats: dict | None = None):
    """Keep list (positions, ascending) of greedy NMS over score-sorted `boxes`, first `max_keep` keeps, computed
    the way the bucketed kernel computes it.  `chunk_sizes`: sizes of the successive chunks (the last repeats)."""
    t = make_thr(thr)
    assert t["fast"], "the bucketed kernel is only selected for screenable thresholds"
    boxes = np.ascontiguousarray(boxes, f32)
    n = boxes.shape[0]
    kept: list[int] = []
    base, ci = 0, 0
    pairs = 0
    while base < n and len(kept) < max_keep:
        kc = chunk_sizes[min(ci, len(chunk_sizes) - 1)]
        ci += 1
        cand = boxes[base:base + kc]
        m = cand.shape[0]
        sa_c = screen_area(cand, t["c2"])
        rng = admissible_ranges(cand, sa_c, t)
        dead = np.zeros(m, bool)
        if kept:
            kb = boxes[np.asarray(kept)]
            sa_k = screen_area(kb, t["c2"])
            cls_k, xb_k, alw_k = keys_of(kb, sa_k)
            adm = _admissible(cls_k, xb_k, alw_k, rng)
            hit = adm & suppress_screen(kb, sa_k, cand, sa_c)
            hit &= suppress_exact(kb, cand, t["up"])
            pairs += int(adm.sum())
            dead = hit.any(axis=0)
        # predecessors among the chunk's survivors (earlier position only)
        cls_c, xb_c, alw_c = keys_of(cand, sa_c)
        adm = _admissible(cls_c, xb_c, alw_c, rng)               # [j (list side), i (walker)]
        adm &= (np.arange(m)[:, None] < np.arange(m)[None, :])
        adm &= ~dead[:, None] & ~dead[None, :]
        pred = adm & suppress_screen(cand, sa_c, cand, sa_c)
        pred &= suppress_exact(cand, cand, t["up"])               # pred[j, i]: j suppresses i
        pairs += int(adm.sum())
        # fix-point: kept once every predecessor is removed; removed as soon as one predecessor is kept
        state = np.where(dead, 2, 0)                              # 0 undecided, 1 kept, 2 removed
        rounds = 0
        while (state == 0).any():
            rounds += 1
            has_kept = (pred & (state == 1)[:, None]).any(axis=0)
            pending = (pred & (state == 0)[:, None]).any(axis=0)
            new = state.copy()
            new[(state == 0) & has_kept] = 2
            new[(state == 0) & ~has_kept & ~pending] = 1
            state = new
        for i in np.nonzero(state == 1)[0]:
            if len(kept) < max_keep:
                kept.append(base + int(i))
        if stats is not None:
            stats.setdefault("rounds", []).append(rounds)
            stats.setdefault("survivors", []).append(int((~dead).sum()))
            stats.setdefault("max_preds", []).append(int(pred.sum(axis=0).max()) if m else 0)
            stats.setdefault("tot_preds", []).append(int(pred.sum()))
            stats.setdefault("kept_after", []).append(len(kept))
        base += m
    if stats is not None:
        stats["pairs"] = pairs
        stats["visited"] = base
    return np.asarray(kept, np.int64)
