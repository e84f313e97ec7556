"""Box sets that stress the bucketed NMS kernels (area-class / x-bin pruning, screening margins, chunk
scheduling).  All boxes are normalised xyxy in [0,1], listed in score order (the first box has the best score).
Used by the GPU parity tests and by the CPU model of the kernel's algorithm (tests/nms_model.py)."""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _clip(b):
    return np.clip(b, 0.0, 1.0).astype(f32)


def rpn_like(seed: int, n: int, hw=(608, 1008)):
    """Decoded RPN proposals of a synthetic head (three anchor scales x three ratios), score sorted."""
    from faster_rcnn_pytorch_b200 import synth
    from oracle import region_oracle as orc
    _, reg, scores = synth.rpn_head_outputs(seed, hw)
    boxes = orc.decode_clip(reg, orc.enumerate_anchors(hw))
    valid = orc.min_size_mask(boxes)
    src = np.nonzero(valid)[0]
    order = orc.sort_desc(scores[valid])[:n]
    return np.ascontiguousarray(boxes[src[order]])


def area_class_boundaries(seed: int, n: int):
    """Areas exactly on the quarter-octave class boundaries 2^e * {1, 1.25, 1.5, 1.75} and one ulp either side;
    widths / heights are dyadic so that x2 - x1 and the product are exact in fp32."""
    rs = np.random.RandomState(seed)
    b = np.zeros((n, 4), np.float64)
    for i in range(n):
        ew, eh = rs.randint(2, 6), rs.randint(2, 6)
        m = (1.0, 1.25, 1.5, 1.75)[rs.randint(4)]
        w, h = 2.0 ** -ew, (2.0 ** -eh) * m
        x1 = rs.randint(0, 64) / 128.0
        y1 = rs.randint(0, 64) / 128.0
        b[i] = (x1, y1, x1 + w, y1 + h)
    b = b.astype(f32)
    # nudge a third of the boxes one ulp in y2: the area moves across the boundary
    k = rs.randint(0, 3, n)
    up = np.nextafter(b[:, 3], f32(2.0)); dn = np.nextafter(b[:, 3], f32(-1.0))
    b[:, 3] = np.where(k == 1, up, np.where(k == 2, dn, b[:, 3]))
    return _clip(b)


def x_bin_edges(seed: int, n: int):
    """x-centres exactly on the bin edges j/8 (and at 0 and 1), widths from tiny to a quarter of the image."""
    rs = np.random.RandomState(seed)
    cx = rs.randint(0, 9, n) / 8.0
    w = 2.0 ** -rs.randint(2, 9, n).astype(np.float64)
    cy = rs.uniform(0.1, 0.9, n)
    h = rs.uniform(0.02, 0.3, n)
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1)
    # duplicates a little to the left / right of the edge: neighbours in adjacent bins that do overlap
    j = rs.randint(0, n, n // 3)
    b[j, 0] += 1e-4; b[j, 2] += 1e-4
    return _clip(b)


def nested_at_threshold(seed: int, n: int, thr: float):
    """Pairs (outer, inner): the inner box is contained in the outer one and its area is thr * (1 +- tiny) of it, so
    IoU = area ratio sits on the decision boundary AND on the edge of the admissible area range."""
    rs = np.random.RandomState(seed)
    m = n // 2
    x1 = rs.uniform(0.0, 0.5, m); y1 = rs.uniform(0.0, 0.5, m)
    w = rs.uniform(0.05, 0.45, m); h = rs.uniform(0.05, 0.45, m)
    outer = np.stack([x1, y1, x1 + w, y1 + h], axis=1)
    eps = rs.choice([-3e-7, -1e-7, 0.0, 1e-7, 3e-7, 1e-3, -1e-3], m)
    f = thr * (1.0 + eps)
    # shrink along x only (left aligned), along y only, or both
    mode = rs.randint(0, 3, m)
    fx = np.where(mode == 0, f, np.where(mode == 1, 1.0, np.sqrt(f)))
    fy = f / fx
    inner = np.stack([x1, y1, x1 + w * fx, y1 + h * fy], axis=1)
    first_outer = rs.randint(0, 2, m).astype(bool)
    a = np.where(first_outer[:, None], outer, inner)
    c = np.where(first_outer[:, None], inner, outer)
    b = np.empty((2 * m, 4))
    b[0::2] = a
    b[1::2] = c
    return _clip(b)


def shifted_at_threshold(seed: int, n: int, thr: float):
    """Pairs at the largest centre distance that still reaches IoU = thr: a box of width w and one of width w / thr
    sharing the left edge (distance w (1 - thr) / (2 thr)), and equal boxes shifted by w (1 - thr) / (1 + thr)."""
    rs = np.random.RandomState(seed)
    m = n // 2
    w = rs.uniform(0.01, 0.3, m) * thr
    h = rs.uniform(0.05, 0.3, m)
    x1 = rs.uniform(0.0, 0.6, m); y1 = rs.uniform(0.0, 0.6, m)
    eps = rs.choice([-3e-7, 0.0, 3e-7, 1e-3, -1e-3], m)
    kind = rs.randint(0, 2, m)
    wide = np.stack([x1, y1, x1 + w / thr * (1 + eps), y1 + h], axis=1)
    d = w * (1 - thr) / (1 + thr) * (1 + eps)
    shift = np.stack([x1 + d, y1, x1 + d + w, y1 + h], axis=1)
    base = np.stack([x1, y1, x1 + w, y1 + h], axis=1)
    other = np.where(kind[:, None] == 0, wide, shift)
    swap = rs.randint(0, 2, m).astype(bool)
    b = np.empty((2 * m, 4))
    b[0::2] = np.where(swap[:, None], other, base)
    b[1::2] = np.where(swap[:, None], base, other)
    return _clip(b)


def x_cut_extremal(seed: int, n: int, thr: float):
    """Pairs (wide box first, narrow box second) with IoU = thr * (1 +- tiny) at the largest x-centre distance a
    suppressing pair can have relative to the NARROW (walking) box, placed so that the wide box's centre lies just
    across an x-bin edge from the point the walker's admissible range reaches: a walk that is a hair too short
    misses the suppressor."""
    rs = np.random.RandomState(seed)
    m = n // 2
    w = rs.uniform(0.02, 0.12, m) * min(1.0, 2.0 * thr)
    h = rs.uniform(0.05, 0.3, m)
    y1 = rs.uniform(0.0, 0.6, m)
    e = rs.randint(2, 7, m) / 8.0                           # the bin edge
    eps = rs.choice([0.0, 1e-7, 3e-7, 1e-6, -1e-7, 1e-4], m)
    wk = w / thr * (1 - eps)                                # wide box: IoU = w / wk = thr (1 + eps)
    right = rs.randint(0, 2, m).astype(bool)                # wide box's centre to the right / left of the narrow one
    tiny = rs.choice([1e-7, 1e-6, 1e-5], m)
    cxk = np.where(right, e + tiny, e - tiny)
    kx1 = cxk - wk / 2
    kx2 = cxk + wk / 2
    cx1 = np.where(right, kx1, kx2 - w)                     # shared left / right edge
    wide = np.stack([kx1, y1, kx2, y1 + h], axis=1)
    narrow = np.stack([cx1, y1, cx1 + w, y1 + h], axis=1)
    b = np.empty((2 * m, 4))
    b[0::2] = wide
    b[1::2] = narrow
    return _clip(b)


def degenerate_mix(seed: int, n: int):
    """Clustered boxes with zero-width, zero-height, inverted and point boxes mixed in (NaN / non-positive areas)."""
    from faster_rcnn_pytorch_b200 import synth
    rs = np.random.RandomState(seed)
    b, _ = synth.random_boxes(seed, n)
    b = b.copy()
    i = rs.permutation(n)
    q = max(n // 40, 1)
    b[i[:q], 2] = b[i[:q], 0]                               # zero width
    b[i[q:2 * q], 3] = b[i[q:2 * q], 1]                     # zero height
    b[i[2 * q:3 * q]] = b[i[2 * q:3 * q]][:, [2, 1, 0, 3]]  # x2 < x1
    b[i[3 * q:4 * q], 2:] = b[i[3 * q:4 * q], :2]           # points
    b[i[4 * q:5 * q]] = b[i[5 * q:6 * q]]                   # exact duplicates
    b[i[6 * q:7 * q], 3] = b[i[6 * q:7 * q], 1] + f32(1e-20)  # denormal-area boxes
    return _clip(b)


def one_class(seed: int, n: int):
    """Every box has the same size (one area class): the area cut removes nothing, dense overlaps."""
    rs = np.random.RandomState(seed)
    c = rs.uniform(0.1, 0.9, (n, 2))
    c[n // 2:] = c[rs.randint(0, n // 2, n - n // 2)] + rs.normal(0, 0.004, (n - n // 2, 2))
    wh = np.array([0.09375, 0.125])
    return _clip(np.concatenate([c - wh / 2, c + wh / 2], axis=1))


def dense_duplicates(seed: int, n: int):
    """A few hundred sites, every site a pile of near-identical boxes: almost everything is suppressed, the
    predecessor lists of the pile members are as long as the pile."""
    rs = np.random.RandomState(seed)
    sites = rs.uniform(0.15, 0.85, (max(n // 60, 1), 2))
    wh = rs.uniform(0.05, 0.25, (sites.shape[0], 2))
    s = rs.randint(0, sites.shape[0], n)
    c = sites[s] + rs.normal(0, 0.0015, (n, 2))
    w = wh[s] * (1 + rs.normal(0, 0.01, (n, 2)))
    return _clip(np.concatenate([c - w / 2, c + w / 2], axis=1))


def staircase(seed: int, n: int):
    """Rows of boxes, each shifted by 1/7 of its width against its left neighbour (IoU 0.75 with the neighbour, 0.56
    with the next one): at thr 0.7 the decision of a box depends on the decision of its neighbour, chains of ~25
    dependent decisions -- the fix-point needs as many rounds.  Scores run along a row, rows are interleaved."""
    rs = np.random.RandomState(seed)
    per = 25
    rows = (n + per - 1) // per
    w, h = 0.2, 0.9 / rows
    r = np.arange(n) % rows
    k = np.arange(n) // rows
    x1 = 0.01 + k * (w / 7.0) + rs.uniform(0, 1e-4, n)
    y1 = 0.02 + r * h
    return _clip(np.stack([x1, y1, x1 + w, y1 + h * 0.9], axis=1))


def sparse(seed: int, n: int):
    """Small boxes that hardly overlap: (almost) everything is kept -- the walk ends by count, not by max_keep."""
    rs = np.random.RandomState(seed)
    c = rs.uniform(0.02, 0.98, (n, 2))
    wh = rs.uniform(0.004, 0.02, (n, 2))
    return _clip(np.concatenate([c - wh / 2, c + wh / 2], axis=1))


CASES = {
    "area_class_boundaries": area_class_boundaries,
    "x_bin_edges": x_bin_edges,
    "degenerate_mix": degenerate_mix,
    "one_class": one_class,
    "dense_duplicates": dense_duplicates,
    "sparse": sparse,
    "staircase": staircase,
}
THR_CASES = {
    "nested_at_threshold": nested_at_threshold,
    "shifted_at_threshold": shifted_at_threshold,
    "x_cut_extremal": x_cut_extremal,
}


def make(name: str, seed: int, n: int, thr: float = 0.7):
    if name == "rpn_like":
        return rpn_like(seed, n)
    if name in THR_CASES:
        return THR_CASES[name](seed, n, thr)
    return CASES[name](seed, n)


ALL = ["rpn_like"] + sorted(CASES) + sorted(THR_CASES)
