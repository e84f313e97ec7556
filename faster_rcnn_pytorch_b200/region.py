"""Batched region-stage pipelines built from the libfrr kernels (the public functional API).

``rpn_proposals`` is the batched equivalent of the reference's ``RegionProposal.forward``
(models/model.py:17-58): decode+clip+min-size -> top-k -> NMS -> first post_nms boxes, for B images
at once, with no host synchronisation (ragged sizes come back as counts on the device).
"""
from __future__ import annotations

import torch

from . import ops

PROPOSAL_MODES = {"train": (12000, 2000), "test": (6000, 300)}   # models/model.py:24-28
RPN_NMS_THRESH = 0.7                                              # models/model.py:53


def rpn_proposals(cls, reg, image_hw=None, mode: str = "train", anchors=None, stride: int = 16, table=None,
                  pre_nms_top_k: int | None = None, post_nms_top_k: int | None = None,
                  nms_thresh: float = RPN_NMS_THRESH, cluster_size: int = 0, return_all: bool = False):
    """cls [B,N,2] logits (or [B,N] fg scores), reg [B,N,4] -> rois [B,post,4] (zero padded), count int32 [B]."""
    pre_k, post_k = PROPOSAL_MODES[mode]
    pre_k = pre_k if pre_nms_top_k is None else int(pre_nms_top_k)
    post_k = post_k if post_nms_top_k is None else int(post_nms_top_k)
    boxes, scores, valid = ops.rpn_decode(reg, cls, image_hw=image_hw, anchors=anchors, stride=stride, table=table)
    k = min(pre_k, boxes.shape[1])
    top = ops.topk_desc(scores, k, valid=valid, boxes=boxes, want_cidx=return_all)
    keep, count, rois = ops.nms_sorted(top["boxes"], nms_thresh, max_keep=post_k, counts=top["count"],
                                       cluster_size=cluster_size)
    if return_all:
        return dict(rois=rois, count=count, keep=keep, topk=top, boxes=boxes, scores=scores, valid=valid)
    return rois, count
