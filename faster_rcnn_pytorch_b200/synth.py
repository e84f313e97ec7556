"""Deterministic synthetic inputs for the region stage.

numpy's legacy ``RandomState`` stream is stable across numpy versions and machines, so the
golden fixtures in ``tests/golden`` store only seeds + outputs for the full-size cases and
every box (build container, GPU box) regenerates bit-identical inputs.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def feature_hw(image_hw, stride: int = 16):
    return int(image_hw[0]) // stride, int(image_hw[1]) // stride


def num_anchors(image_hw, stride: int = 16, per_cell: int = 9) -> int:
    fh, fw = feature_hw(image_hw, stride)
    return fh * fw * per_cell


def unique_scores(rs: np.random.RandomState, n: int) -> np.ndarray:
    """fp32 objectness scores in (0,1) with no two equal (top-k parity is defined on tie-free
    keys: the reference's ``sort(descending=True)`` tie order is unspecified)."""
    z = rs.standard_normal((n, 2)).astype(f32)
    s = (1.0 / (1.0 + np.exp(-(z[:, 1].astype(np.float64) - z[:, 0].astype(np.float64))))).astype(f32)
    for _ in range(64):
        _, first = np.unique(s, return_index=True)
        dup = np.ones(n, dtype=bool)
        dup[first] = False
        nd = int(dup.sum())
        if nd == 0:
            break
        s[dup] = rs.uniform(0.0, 1.0, nd).astype(f32)
    assert np.unique(s).shape[0] == n
    return s


def rpn_head_outputs(seed: int, image_hw, with_logits: bool = True):
    """cls logits [N,2] ~ N(0,1), reg [N,4] ~ N(0,0.2^2), tie-free fg scores [N]."""
    n = num_anchors(image_hw)
    rs = np.random.RandomState(seed)
    logits = rs.standard_normal((n, 2)).astype(f32)
    reg = (rs.standard_normal((n, 4)) * 0.2).astype(f32)
    scores = unique_scores(rs, n)
    return logits, reg, scores


def gt_boxes(seed: int, g: int = 8, num_classes: int = 20):
    """GT boxes xy1 ~ U(0,0.7), wh ~ U(0.05,0.3) (normalised xyxy), labels randint(0,num_classes)."""
    rs = np.random.RandomState(seed)
    xy1 = rs.uniform(0.0, 0.7, (g, 2))
    wh = rs.uniform(0.05, 0.3, (g, 2))
    boxes = np.concatenate([xy1, xy1 + wh], axis=1).astype(f32)
    labels = rs.randint(0, num_classes, g).astype(np.int64)
    return boxes, labels


def features(seed: int, batch: int, channels: int, fh: int, fw: int) -> np.ndarray:
    rs = np.random.RandomState(seed)
    return rs.standard_normal((batch, channels, fh, fw)).astype(f32)


def random_rois(seed: int, k: int, fh: int, fw: int, batch: int = 1) -> np.ndarray:
    """[K,5] (batch idx, x1,y1,x2,y2) in feature-map coordinates, some touching/crossing borders."""
    rs = np.random.RandomState(seed)
    x1 = rs.uniform(-1.0, fw - 1.0, k)
    y1 = rs.uniform(-1.0, fh - 1.0, k)
    w = rs.uniform(0.2, fw * 0.7, k)
    h = rs.uniform(0.2, fh * 0.7, k)
    b = rs.randint(0, batch, k)
    return np.stack([b, x1, y1, np.minimum(x1 + w, fw + 0.5), np.minimum(y1 + h, fh + 0.5)], axis=1).astype(f32)


def head_outputs(seed: int, r: int, num_classes: int):
    """Fast R-CNN head outputs for the predict tail: cls [R,C] ~ N(0,1), reg [R,4C] ~ N(0,1)."""
    rs = np.random.RandomState(seed)
    cls = rs.standard_normal((r, num_classes)).astype(f32)
    reg = rs.standard_normal((r, 4 * num_classes)).astype(f32)
    return cls, reg


def random_boxes(seed: int, n: int, cluster: bool = True):
    """Normalised xyxy boxes with heavy overlap (clustered centres) + tie-free scores."""
    rs = np.random.RandomState(seed)
    if cluster:
        centres = rs.uniform(0.1, 0.9, (max(n // 40, 1), 2))
        c = centres[rs.randint(0, centres.shape[0], n)] + rs.normal(0, 0.02, (n, 2))
    else:
        c = rs.uniform(0.0, 1.0, (n, 2))
    wh = rs.uniform(0.02, 0.3, (n, 2))
    b = np.concatenate([c - wh / 2, c + wh / 2], axis=1)
    b = np.clip(b, 0.0, 1.0).astype(f32)
    s = unique_scores(rs, n)
    return b, s


def pyramid_inputs(seed=8100, B=2, C=8, image_hw=(256, 320), K=80):
    """Seeded pyramid features (strides 4..32) and rois in image coordinates spanning all levels."""
    rs = np.random.RandomState(seed)
    feats = [rs.standard_normal((B, C, image_hw[0] // s, image_hw[1] // s)).astype(np.float32) for s in (4, 8, 16, 32)]
    side = np.exp(rs.uniform(np.log(8), np.log(300), K))
    ar = np.exp(rs.uniform(-0.7, 0.7, K))
    w, h = side * np.sqrt(ar), side / np.sqrt(ar)
    cx, cy = rs.uniform(0, image_hw[1], K), rs.uniform(0, image_hw[0], K)
    x1, y1 = np.clip(cx - w / 2, 0, image_hw[1] - 2), np.clip(cy - h / 2, 0, image_hw[0] - 2)
    x2, y2 = np.clip(cx + w / 2, x1 + 1, image_hw[1]), np.clip(cy + h / 2, y1 + 1, image_hw[0])
    b = rs.randint(0, B, K)
    rois5 = np.stack([b, x1, y1, x2, y2], axis=1).astype(np.float32)
    rois5[0, 1:] = [10, 10, 10 + 112, 10 + 112]       # sqrt(area) = 112 = 224 / 2: exactly on a level boundary
    rois5[1, 1:] = [0, 0, 224, 224]
    return feats, rois5


def loss_inputs(seed=8300, N=20646, S=128, C=21, n_pos=40, n_neg=216, frc_pos=32):
    """Seeded predictions / targets with the shapes of losses/loss.py:24-59: RPN labels with n_pos ones, n_neg zeros and
    -1 elsewhere; Fast R-CNN classes with frc_pos foreground rows first (like models/model.py:165); some regression
    differences fall on either side of the SmoothL1 beta (1/9 and 1)."""
    rs = np.random.RandomState(seed)
    f32 = np.float32
    rpn_cls = rs.standard_normal((N, 2)).astype(f32)
    rpn_reg = (rs.standard_normal((N, 4)) * 0.3).astype(f32)
    t = np.full((N,), -1, np.int64)
    sel = rs.permutation(N)[:n_pos + n_neg]
    t[sel[:n_pos]] = 1
    t[sel[n_pos:]] = 0
    rpn_treg = (rs.standard_normal((N, 4)) * 0.3).astype(f32)
    frc_cls = rs.standard_normal((S, C)).astype(f32) * 2
    frc_reg = rs.standard_normal((S, C, 4)).astype(f32)
    c = np.zeros((S,), np.int64)
    c[:frc_pos] = rs.randint(1, C, frc_pos)
    frc_treg = rs.standard_normal((S, 4)).astype(f32)
    frc_treg[0] = frc_reg[0, c[0]]            # a zero difference (gradient 0 at the kink of |x|)
    return dict(rpn_cls=rpn_cls, rpn_reg=rpn_reg, rpn_tcls=t, rpn_treg=rpn_treg, frc_cls=frc_cls, frc_reg=frc_reg,
                frc_tcls=c, frc_treg=frc_treg)
