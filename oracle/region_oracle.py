"""CPU oracle for the Faster R-CNN region stage  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy (+ small C helper, see ``oracle/c/region_oracle.c``) restatement of
the reference's region-stage algorithm.  It is the *checker* for the CUDA path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``faster_rcnn_pytorch_b200/``
imports it, and the product path raises when the CUDA library is missing.

Parity pin: the reference has no tests and pins no torchvision version, so
the pins are (a) golden vectors produced by importing the real reference from
``/root/reference`` plus torchvision 0.26 CPU kernels in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``) and (b) the known-answer tests
KAT-1..7 of SURVEY.md §8c.  ``tests/test_oracle_golden.py`` checks every function here against
those.

All arithmetic is IEEE fp32 with the operation order of the reference; each function cites
the reference lines it follows (paths relative to /root/reference, ``TV:`` = torchvision).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

# ----------------------------------------------------------------------------------------
# box parametrisation helpers  (utils/util.py:15-50)
# ----------------------------------------------------------------------------------------


def corners_to_center(xy: np.ndarray) -> np.ndarray:
    """utils/util.py:22-26  ``xy_to_cxcy``: (x1,y1,x2,y2) -> ((x2+x1)/2, (y2+y1)/2, x2-x1, y2-y1)."""
    xy = np.asarray(xy, dtype=f32)
    lo, hi = xy[..., :2], xy[..., 2:]
    return np.concatenate([(hi + lo) / f32(2), hi - lo], axis=-1).astype(f32)


def center_to_corners(c: np.ndarray) -> np.ndarray:
    """utils/util.py:15-19  ``cxcy_to_xy``: (cx,cy,w,h) -> (c - wh/2, c + wh/2)."""
    c = np.asarray(c, dtype=f32)
    half = c[..., 2:] / f32(2)
    return np.concatenate([c[..., :2] - half, c[..., :2] + half], axis=-1).astype(f32)


def encode_boxes(gt_c: np.ndarray, anc_c: np.ndarray) -> np.ndarray:
    """utils/util.py:39-43  ``encode``: ((g_c - a_c)/a_wh, log(g_wh/a_wh)), centre form in/out."""
    gt_c = np.asarray(gt_c, dtype=f32)
    anc_c = np.asarray(anc_c, dtype=f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_c = (gt_c[:, :2] - anc_c[:, :2]) / anc_c[:, 2:]
        t_wh = np.log(gt_c[:, 2:] / anc_c[:, 2:])
    return np.concatenate([t_c, t_wh], axis=1).astype(f32)


def decode_boxes(t: np.ndarray, anc_c: np.ndarray) -> np.ndarray:
    """utils/util.py:46-50  ``decode``: (t_xy*a_wh + a_c, exp(t_wh)*a_wh), centre form in/out."""
    t = np.asarray(t, dtype=f32)
    anc_c = np.asarray(anc_c, dtype=f32)
    c = t[:, :2] * anc_c[:, 2:] + anc_c[:, :2]
    wh = np.exp(t[:, 2:]) * anc_c[:, 2:]
    return np.concatenate([c, wh], axis=1).astype(f32)


def pairwise_iou_eps(a: np.ndarray, b: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """utils/util.py:66-102  ``find_jaccard_overlap`` / ``find_intersection``.

    inter = clamp(min(x2)-max(x1),0) * clamp(min(y2)-max(y1),0)
    union = ((area_a + area_b) - inter) + eps_f32 ; iou = inter / union      (all fp32)
    """
    a = np.asarray(a, dtype=f32)
    b = np.asarray(b, dtype=f32)
    lo = np.maximum(a[:, None, :2], b[None, :, :2])
    hi = np.minimum(a[:, None, 2:], b[None, :, 2:])
    d = np.maximum(hi - lo, f32(0))
    inter = d[..., 0] * d[..., 1]
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    union = ((area_a[:, None] + area_b[None, :]) - inter) + f32(eps)
    return (inter / union).astype(f32)


# ----------------------------------------------------------------------------------------
# anchors  (anchor.py:15-55)
# ----------------------------------------------------------------------------------------


def anchor_base_table(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)) -> np.ndarray:
    """anchor.py:15-32  nine base boxes, ratio-major; float64 math stored to fp32."""
    tab = np.zeros((len(ratios) * len(scales), 4), dtype=f32)
    c = base_size / 2.0
    for i, r in enumerate(ratios):
        for j, s in enumerate(scales):
            w = base_size * s * np.sqrt(r)
            h = base_size * s * np.sqrt(1.0 / r)
            tab[i * len(scales) + j] = (c - w / 2.0, c - h / 2.0, c + w / 2.0, c + h / 2.0)
    return tab


def enumerate_anchors(image_hw, base_size=16, table: np.ndarray | None = None) -> np.ndarray:
    """anchor.py:34-55  anchor[(y*fw + x)*A + a] = fl32(base[a] + 16*(x,y,x,y)) / (W,H,W,H).

    The reference adds an int64 shift to the fp32 table (numpy promotes to float64; the sum
    is exact there) and rounds once to fp32, then divides in place by an int64 divisor
    (float64 loop, rounded to fp32).  Both equal the correctly rounded fp32 add / fp32
    divide, which is what is written here and what the CUDA kernel does with
    ``__fadd_rn`` / ``__fdiv_rn``.
    """
    H, W = int(image_hw[0]), int(image_hw[1])
    fh, fw = H // base_size, W // base_size
    tab = anchor_base_table(base_size) if table is None else np.asarray(table, dtype=f32)
    A = tab.shape[0]
    xs = (np.arange(fw, dtype=np.int64) * base_size).astype(f32)
    ys = (np.arange(fh, dtype=np.int64) * base_size).astype(f32)
    out = np.empty((fh, fw, A, 4), dtype=f32)
    out[..., 0] = tab[None, None, :, 0] + xs[None, :, None]
    out[..., 1] = tab[None, None, :, 1] + ys[:, None, None]
    out[..., 2] = tab[None, None, :, 2] + xs[None, :, None]
    out[..., 3] = tab[None, None, :, 3] + ys[:, None, None]
    out[..., 0::2] /= f32(W)
    out[..., 1::2] /= f32(H)
    return out.reshape(fh * fw * A, 4)


def tv_anchor_base(size: float, aspect_ratios=(0.5, 1.0, 2.0)) -> np.ndarray:
    """TV models/detection/anchor_utils.py ``AnchorGenerator.generate_anchors`` for one size (the configuration of
    models/new_model.py:23-25): all fp32, ``torch.round`` = half to even."""
    r = np.asarray(aspect_ratios, dtype=f32)
    hr = np.sqrt(r).astype(f32)
    wr = (f32(1) / hr).astype(f32)
    ws = (wr * f32(size)).astype(f32)
    hs = (hr * f32(size)).astype(f32)
    return np.rint((np.stack([-ws, -hs, ws, hs], axis=1) / f32(2)).astype(f32)).astype(f32)


def tv_anchors_pyramid(feature_hws, image_hw, sizes=(32, 64, 128, 256, 512), aspect_ratios=(0.5, 1.0, 2.0)) -> np.ndarray:
    """models/new_model.py:43-44: ``AnchorGenerator(...)(ImageList(x, [(w, h)]), features)[0] / (w, h, w, h)``:
    per level strides = image // grid (integer), shifts (x, y, x, y) row-major over the grid, anchors = shift + base
    (cell-major, anchor-minor), levels concatenated, fp32 division by the image size."""
    H, W = int(image_hw[0]), int(image_hw[1])
    out = []
    for (fh, fw), sz in zip(feature_hws, sizes):
        base = tv_anchor_base(sz, aspect_ratios)
        sy, sx = H // int(fh), W // int(fw)
        xs = (np.arange(fw, dtype=np.int32) * sx).astype(f32)
        ys = (np.arange(fh, dtype=np.int32) * sy).astype(f32)
        yy, xx = np.meshgrid(ys, xs, indexing="ij")
        sh = np.stack([xx.ravel(), yy.ravel(), xx.ravel(), yy.ravel()], axis=1).astype(f32)
        out.append((sh[:, None, :] + base[None, :, :]).astype(f32).reshape(-1, 4))
    a = np.concatenate(out, axis=0)
    return (a / np.array([W, H, W, H], dtype=f32)).astype(f32)


# ----------------------------------------------------------------------------------------
# NMS  (TV: csrc/ops/cpu/nms_kernel.cpp semantics)
# ----------------------------------------------------------------------------------------


def nms(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """Greedy NMS, torchvision CPU semantics (call sites models/model.py:53, :394).

    order = stable descending sort of scores; IoU = inter / ((area_i + area_j) - inter) in
    fp32; box j is suppressed when ``(double)iou > iou_threshold``.  Returns int64 indices
    into ``boxes`` in descending-score order.  Uses the C helper when built, else a
    vectorised numpy sweep (same arithmetic).
    """
    boxes = np.ascontiguousarray(boxes, dtype=f32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=f32).reshape(-1)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-scores.astype(np.float64), kind="stable").astype(np.int64)
    # -scores in float64 is exact, stable argsort => descending, ties by lower index
    from . import _cbridge

    lib = _cbridge.load()
    if lib is not None:
        return _cbridge.nms_sorted(lib, boxes, order, float(iou_threshold))
    return _nms_numpy(boxes, order, float(iou_threshold))


def _nms_numpy(boxes: np.ndarray, order: np.ndarray, thr: float) -> np.ndarray:
    b = boxes[order]
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    area = (x2 - x1) * (y2 - y1)
    n = b.shape[0]
    dead = np.zeros(n, dtype=bool)
    keep = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(n):
            if dead[i]:
                continue
            keep.append(i)
            if i + 1 == n:
                break
            w = np.maximum(f32(0), np.minimum(x2[i], x2[i + 1:]) - np.maximum(x1[i], x1[i + 1:]))
            h = np.maximum(f32(0), np.minimum(y2[i], y2[i + 1:]) - np.maximum(y1[i], y1[i + 1:]))
            inter = w * h
            ovr = inter / ((area[i] + area[i + 1:]) - inter)
            dead[i + 1:] |= ovr.astype(np.float64) > thr
    return order[np.asarray(keep, dtype=np.int64)]


# ----------------------------------------------------------------------------------------
# RPN proposal layer  (models/model.py:17-58)
# ----------------------------------------------------------------------------------------

PROPOSAL_MODES = {"train": (12000, 2000), "test": (6000, 300)}


def fg_softmax(cls_logits: np.ndarray) -> np.ndarray:
    """models/model.py:20  softmax over the 2 logits, foreground column."""
    z = np.asarray(cls_logits, dtype=f32)
    m = z.max(axis=-1, keepdims=True)
    e = np.exp(z - m)
    return (e[..., 1] / (e[..., 0] + e[..., 1])).astype(f32)


def decode_clip(reg: np.ndarray, anchor: np.ndarray) -> np.ndarray:
    """models/model.py:31-34  decode against centre-form anchors, back to corners, clamp [0,1]."""
    roi = center_to_corners(decode_boxes(reg, corners_to_center(anchor)))
    return np.clip(roi, f32(0), f32(1)).astype(f32)


def min_size_mask(roi: np.ndarray, min_size: float = 1.0) -> np.ndarray:
    """models/model.py:37-39  keep boxes with h >= 0.001f and w >= 0.001f (fp32 compare)."""
    ws = roi[:, 2] - roi[:, 0]
    hs = roi[:, 3] - roi[:, 1]
    t = f32(min_size / 1000)
    return (hs >= t) & (ws >= t)


def sort_desc(scores: np.ndarray) -> np.ndarray:
    """models/model.py:44  descending sort; ties resolved lower-index-first (the reference's
    tie order is unspecified; parity is defined on tie-free keys)."""
    return np.argsort(-np.asarray(scores, dtype=f32).astype(np.float64), kind="stable").astype(np.int64)


def region_proposal(cls_logits, reg, anchor, mode: str = "train", scores=None, pre_k=None, post_k=None, min_size: float = 1.0):
    """models/model.py:17-58  full proposal layer for ONE image.

    Returns a dict with every intermediate the CUDA path is checked against:
    ``score`` [N], ``boxes`` [N,4] (decoded+clipped), ``valid`` [N] bool,
    ``topk_idx`` (indices into the compacted array, as the reference), ``topk_src``
    (indices into the full N), ``topk_boxes``, ``topk_scores``, ``keep`` (positions in the
    top-k list) and ``rois``.
    ``scores`` overrides the softmax (used to feed identical fp32 scores to both sides).
    """
    if pre_k is None or post_k is None:
        pre_k, post_k = PROPOSAL_MODES[mode]        # models/new_model.py:52-56 passes 4000|2000 -> 1000, min_size 10
    score = fg_softmax(cls_logits) if scores is None else np.asarray(scores, dtype=f32)
    boxes = decode_clip(reg, anchor)
    valid = min_size_mask(boxes, min_size)
    comp_boxes = boxes[valid]
    comp_score = score[valid]
    src = np.nonzero(valid)[0].astype(np.int64)
    order = sort_desc(comp_score)
    k = min(pre_k, order.shape[0])
    top = order[:k]
    tb = comp_boxes[top]
    ts = comp_score[top]
    keep = nms(tb, ts, 0.7)[:post_k]
    return dict(score=score, boxes=boxes, valid=valid, topk_idx=top, topk_src=src[top],
                topk_boxes=tb, topk_scores=ts, keep=keep, rois=tb[keep])


# ----------------------------------------------------------------------------------------
# host RNG: torch.randperm on the CPU generator  (ATen randperm_cpu; SURVEY KAT-5)
# ----------------------------------------------------------------------------------------


class HostRandperm:
    """mt19937 (``init_genrand(seed)``) + forward Fisher-Yates ``z = next_u32() % (n-i)``.

    Restates what ``torch.manual_seed(seed); torch.randperm(n)`` draws on the CPU generator
    (models/model.py:149,155,228,235 call it that way).  ``randperm(0)`` / ``randperm(1)``
    consume no random numbers.
    """

    def __init__(self, seed: int):
        self._bg = np.random.MT19937()
        self._bg._legacy_seeding(int(seed) & 0xFFFFFFFF)

    def __call__(self, n: int) -> np.ndarray:
        r = np.arange(n, dtype=np.int64)
        if n < 2:
            return r
        raw = self._bg.random_raw(n - 1)
        for i in range(n - 1):
            z = int(raw[i]) % (n - i)
            r[i], r[z + i] = r[z + i], r[i]
        return r


# ----------------------------------------------------------------------------------------
# target makers  (models/model.py:123-266)
# ----------------------------------------------------------------------------------------


def rpn_targets(gt: np.ndarray, anchor: np.ndarray, randperm, variant: str = "vgg") -> dict:
    """models/model.py:186-266  ``RPNTargetMaker.forward`` for one image; ``variant="fpn"`` restates
    models/new_model.py:299-349 instead (box_iou without eps, util/box_ops.py:24-37; no inside-image filter;
    every anchor whose IoU equals a GT's best IoU becomes positive, :316-318).

    ``randperm`` is a callable n -> int64 permutation (torch.randperm or HostRandperm).
    Returns labels int64 [N] in {-1,0,1}, reg fp32 [N,4], and the intermediates
    (``inside``, ``iou_max``, ``argmax``, ``gt_argmax``, ``n_pos``, ``n_neg``).
    """
    gt = np.asarray(gt, dtype=f32).reshape(-1, 4)
    anchor = np.asarray(anchor, dtype=f32)
    if gt.shape[0] == 0:
        raise IndexError("max(): Expected reduction dim 1 to have non-zero size")  # :199
    fpn = variant == "fpn"
    inside = (anchor[:, 0] >= 0) & (anchor[:, 1] >= 0) & (anchor[:, 2] <= 1) & (anchor[:, 3] <= 1)
    if fpn:
        inside = np.ones(anchor.shape[0], dtype=bool)
    a_in = anchor[inside]
    iou = pairwise_iou_eps(a_in, gt, 0.0 if fpn else 1e-5)  # :198 / new_model.py:309
    iou_max = iou.max(axis=1)
    argmax = iou.argmax(axis=1).astype(np.int64)           # first index on ties
    gt_argmax = iou.argmax(axis=0).astype(np.int64)        # :206, first index on ties
    label = np.full(a_in.shape[0], -1, dtype=np.int64)
    label[iou_max < f32(0.3)] = 0                          # :202 fp32 compare
    if fpn:
        label[(iou == iou.max(axis=0)[None, :]).any(axis=1)] = 1   # new_model.py:316-318, ties included
    else:
        label[gt_argmax] = 1                               # :213
    label[iou_max >= f32(0.7)] = 1                         # :216
    n_pos = int((label == 1).sum())
    n_neg = int((label == 0).sum())
    n_pos0, n_neg0 = n_pos, n_neg
    if n_pos > 128:                                        # :225-229
        pos_idx = np.nonzero(label == 1)[0]
        perm = np.asarray(randperm(pos_idx.shape[0]))
        label[pos_idx[perm[128:]]] = -1
    if n_neg > 256 - n_pos:                                # :231-236
        if n_pos > 128:
            n_pos = 128
        neg_idx = np.nonzero(label == 0)[0]
        perm = np.asarray(randperm(neg_idx.shape[0]))
        label[neg_idx[perm[256 - n_pos:]]] = -1
    reg_in = encode_boxes(corners_to_center(gt[argmax]), corners_to_center(a_in))   # :253
    N = anchor.shape[0]
    labels = np.full(N, -1, dtype=np.int64)
    labels[inside] = label
    reg = np.zeros((N, 4), dtype=f32)
    reg[inside] = reg_in
    full_argmax = np.zeros(N, dtype=np.int64)
    full_argmax[inside] = argmax
    full_max = np.zeros(N, dtype=f32)
    full_max[inside] = iou_max
    return dict(labels=labels, reg=reg, inside=inside, iou_max=full_max, argmax=full_argmax,
                gt_argmax=np.nonzero(inside)[0][gt_argmax], n_pos=n_pos0, n_neg=n_neg0)


def frcnn_targets(gt: np.ndarray, gt_label: np.ndarray, rois: np.ndarray, randperm, variant: str = "vgg") -> dict:
    """models/model.py:127-179  ``FastRcnnTargetMaker.forward`` for one image; ``variant="fpn"`` restates
    models/new_model.py:153-206 (box_iou without eps, 512 samples with at most 128 positives).

    rois <- cat(rois, gt); IoU (with eps) vs gt; row max/argmax; cls = label[argmax]+1;
    n_pos = min(#(IoU>=0.5f), 32); pos/neg sampled with two randperm draws (always drawn);
    reg = encode(gt[argmax][keep], rois[keep]) / (0.1,0.1,0.2,0.2).
    """
    gt = np.asarray(gt, dtype=f32).reshape(-1, 4)
    rois = np.concatenate([np.asarray(rois, dtype=f32).reshape(-1, 4), gt], axis=0)      # :135
    fpn = variant == "fpn"
    batch, max_pos = (512, 128) if fpn else (128, 32)
    iou = pairwise_iou_eps(rois, gt, 0.0 if fpn else 1e-5)
    iou_max = iou.max(axis=1)
    argmax = iou.argmax(axis=1).astype(np.int64)
    # :141 adds 1 (labels are 0-based there); new_model.py:166 uses the labels as they are
    cls_all = np.asarray(gt_label)[argmax] + (0 if fpn else 1)
    is_pos = iou_max >= f32(0.5)
    n_pos = int(min(int(is_pos.sum()), max_pos))                                         # :144
    pos_idx = np.nonzero(is_pos)[0]
    perm = np.asarray(randperm(pos_idx.shape[0]))
    pos_idx = pos_idx[perm[:n_pos]]
    n_neg = batch - n_pos
    neg_idx = np.nonzero((iou_max < f32(0.5)) & (iou_max >= f32(0.0)))[0]                # :153
    perm = np.asarray(randperm(neg_idx.shape[0]))
    neg_idx = neg_idx[perm[:n_neg]]
    keep = np.concatenate([pos_idx, neg_idx]).astype(np.int64)
    cls = cls_all[keep].copy()
    cls[n_pos:] = 0                                                                      # :165
    cls = cls.astype(np.int64)
    sample_rois = rois[keep]
    reg = encode_boxes(corners_to_center(gt[argmax][keep]), corners_to_center(sample_rois))
    reg = (reg - f32(0.0)) / np.array([0.1, 0.1, 0.2, 0.2], dtype=f32)                   # :174-177
    return dict(cls=cls, reg=reg.astype(f32), sample_rois=sample_rois, keep=keep,
                iou_max=iou_max, argmax=argmax, n_pos=n_pos)


# ----------------------------------------------------------------------------------------
# RoIPool / RoIAlign  (TV CPU kernels; call sites models/model.py:97,113, new_model.py:127,143)
# ----------------------------------------------------------------------------------------


def roi_pool_forward(feat: np.ndarray, rois5: np.ndarray, pooled=(7, 7), spatial_scale=1.0):
    """TV ``roi_pool_forward_kernel_impl``.  feat [B,C,H,W] fp32, rois5 [K,5]
    (batch idx, x1,y1,x2,y2).  Returns (out [K,C,ph,pw] fp32, argmax int32 same shape)."""
    from . import _cbridge

    lib = _cbridge.load(required=True)
    return _cbridge.roi_pool_fwd(lib, feat, rois5, pooled, float(spatial_scale))


def roi_pool_backward(grad_out, argmax, rois5, feat_shape):
    """TV ``roi_pool_backward_kernel_impl``: grad_in[b,c,argmax] += grad_out."""
    from . import _cbridge

    lib = _cbridge.load(required=True)
    return _cbridge.roi_pool_bwd(lib, grad_out, argmax, rois5, feat_shape)


def roi_align_forward(feat, rois5, pooled=(7, 7), spatial_scale=1.0, sampling_ratio=2, aligned=False):
    """TV ``roi_align_forward_kernel_impl``."""
    from . import _cbridge

    lib = _cbridge.load(required=True)
    return _cbridge.roi_align_fwd(lib, feat, rois5, pooled, float(spatial_scale), int(sampling_ratio), bool(aligned))


def roi_align_backward(grad_out, rois5, feat_shape, pooled=(7, 7), spatial_scale=1.0, sampling_ratio=2,
                       aligned=False):
    """TV ``roi_align_backward_kernel_impl``."""
    from . import _cbridge

    lib = _cbridge.load(required=True)
    return _cbridge.roi_align_bwd(lib, grad_out, rois5, feat_shape, pooled, float(spatial_scale),
                                  int(sampling_ratio), bool(aligned))


def fpn_levels(rois5: np.ndarray, k_min: int, k_max: int, canonical_scale: float = 224.0, canonical_level: int = 4):
    """TV ops/poolers.py ``LevelMapper.__call__`` (used by MultiScaleRoIAlign, models/new_model.py:127,143):
    floor(lvl0 + log2(sqrt(area) / s0) + 1e-6) clamped to [k_min, k_max], minus k_min; fp32 throughout."""
    r = np.asarray(rois5, dtype=f32)
    area = (r[:, 3] - r[:, 1]) * (r[:, 4] - r[:, 2])
    s = np.sqrt(area).astype(f32)
    t = np.floor((f32(canonical_level) + np.log2(s / f32(canonical_scale)).astype(f32)) + f32(1e-6))
    return (np.clip(t, k_min, k_max).astype(np.int64) - k_min)


def infer_scales(feature_shapes, image_shapes):
    """TV ops/poolers.py ``_setup_scales``: 2 ** round(log2(feat_dim0 / max image_dim0)) per level -> (scales, k_min, k_max)."""
    import math
    d0 = max(int(sh[0]) for sh in image_shapes)
    scales = [2.0 ** float(round(math.log2(float(fs[-2]) / float(d0)))) for fs in feature_shapes]
    return scales, int(-math.log2(scales[0])), int(-math.log2(scales[-1]))


def multiscale_roi_align(features, rois5, image_shapes, pooled=(7, 7), sampling_ratio=2):
    """TV ops/poolers.py ``_multiscale_roi_align``: every roi is pooled from the pyramid level LevelMapper assigns."""
    scales, k_min, k_max = infer_scales([f.shape for f in features], image_shapes)
    rois5 = np.asarray(rois5, dtype=f32)
    if len(features) == 1:
        return roi_align_forward(features[0], rois5, pooled, scales[0], sampling_ratio, False), np.zeros(len(rois5), np.int64)
    lv = fpn_levels(rois5, k_min, k_max)
    out = np.zeros((rois5.shape[0], features[0].shape[1]) + tuple(pooled), dtype=f32)
    for l, (f, sc) in enumerate(zip(features, scales)):
        idx = np.nonzero(lv == l)[0]
        if len(idx):
            out[idx] = roi_align_forward(f, rois5[idx], pooled, sc, sampling_ratio, False)
    return out, lv


def smooth_l1(pred, target, beta):
    """losses/loss.py:5-14."""
    x = np.abs(np.asarray(pred, f32) - np.asarray(target, f32))
    return np.where(x >= f32(beta), x - f32(0.5) * f32(beta), f32(0.5) * x * x / f32(beta)).astype(f32)


def _cross_entropy_rows(logits, target):
    l = np.asarray(logits, np.float64)
    m = l.max(axis=1, keepdims=True)
    lse = (m + np.log(np.exp(l - m).sum(axis=1, keepdims=True)))[:, 0]
    return lse - l[np.arange(len(l)), target]


def region_loss(rpn_cls, rpn_reg, rpn_tcls, rpn_treg, frc_cls, frc_reg, frc_tcls, frc_treg):
    """losses/loss.py:17-82 + the class-row gather of models/model.py:340-341 for one image ->
    (total, rpn_cls, rpn_reg, frcnn_cls, frcnn_reg).  frc_reg [S,C,4] (head output) or [S,4] (already gathered);
    negative Fast R-CNN classes mark padding rows of a short sample.  Sums are accumulated in float64 (the checker)."""
    t = np.asarray(rpn_tcls)
    valid, pos = t >= 0, t > 0
    nv = float(valid.sum())
    l_rc = _cross_entropy_rows(np.asarray(rpn_cls)[valid], t[valid]).sum() / nv if nv else float("nan")      # :32
    l_rr = smooth_l1(np.asarray(rpn_reg)[pos], np.asarray(rpn_treg)[pos], 1 / 9).astype(np.float64).sum() / nv if nv else float("nan")
    c = np.asarray(frc_tcls)
    use, fpos = c >= 0, c > 0
    nf = float(use.sum())
    l_fc = _cross_entropy_rows(np.asarray(frc_cls)[use], c[use]).sum() / nf                                   # :55
    reg = np.asarray(frc_reg)
    rows = reg[np.arange(len(c)), np.maximum(c, 0)] if reg.ndim == 3 else reg                                # model.py:340-341
    l_fr = smooth_l1(rows[fpos], np.asarray(frc_treg)[fpos], 1.0).astype(np.float64).sum() / nf              # :56-59
    return np.asarray([l_rc + l_rr + l_fc + l_fr, l_rc, l_rr, l_fc, l_fr], np.float64)


def scale_rois(rois: np.ndarray, fh: int, fw: int, batch_index: int = 0) -> np.ndarray:
    """models/model.py:104-110 + TV ops/_utils.py:18-25: roi*[fw,fh,fw,fh], prepend batch idx."""
    r = np.asarray(rois, dtype=f32) * np.array([fw, fh, fw, fh], dtype=f32)
    return np.concatenate([np.full((r.shape[0], 1), batch_index, dtype=f32), r], axis=1)


# ----------------------------------------------------------------------------------------
# detection post-processing  (models/model.py:369-402)
# ----------------------------------------------------------------------------------------


def softmax_rows(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=f32)
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return (e / e.sum(axis=-1, keepdims=True, dtype=f32)).astype(f32)


def decode_classwise(cls_logits, reg, rois, num_classes, prob=None):
    """models/model.py:369-378  softmax; reg*(0.1,0.1,0.2,0.2); per-class decode; clamp [0,1]."""
    p = softmax_rows(cls_logits) if prob is None else np.asarray(prob, dtype=f32)
    r = np.asarray(reg, dtype=f32).reshape(-1, num_classes, 4) * np.array([0.1, 0.1, 0.2, 0.2], dtype=f32)
    R = r.shape[0]
    rr = np.broadcast_to(np.asarray(rois, dtype=f32).reshape(R, 1, 4), r.shape).reshape(-1, 4)
    b = center_to_corners(decode_boxes(r.reshape(-1, 4), corners_to_center(rr)))
    return p, np.clip(b.reshape(R, num_classes * 4), f32(0), f32(1)).astype(f32)


def suppress(raw_cls_bbox, raw_prob, num_classes, thres=0.05, iou_thr=0.3):
    """models/model.py:382-402  per-class: prob > thres (fp32 compare of a python float cast
    to fp32), nms @0.3, class-major concat, labels l-1 int32."""
    bb = np.asarray(raw_cls_bbox, dtype=f32).reshape(-1, num_classes, 4)
    pr = np.asarray(raw_prob, dtype=f32)
    out_b, out_l, out_s = [], [], []
    for l in range(1, num_classes):
        m = pr[:, l] > f32(thres)
        b = bb[m, l, :]
        s = pr[m, l]
        k = nms(b, s, iou_thr)
        out_b.append(b[k])
        out_l.append(np.full(len(k), l - 1, dtype=np.int32))
        out_s.append(s[k])
    return (np.concatenate(out_b, axis=0).astype(f32).reshape(-1, 4),
            np.concatenate(out_l, axis=0).astype(np.int32),
            np.concatenate(out_s, axis=0).astype(f32))
