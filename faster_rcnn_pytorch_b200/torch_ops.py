"""``torch.ops.frr.*``: the region-stage kernels registered as PyTorch custom ops (``torch.library.custom_op`` with fake
/ meta implementations and autograd formulas), with the schemas of the torchvision ops the reference calls
(models/model.py:6-9,53,97,113,394; models/new_model.py:127,143):

    frr::nms(Tensor dets, Tensor scores, float iou_threshold) -> Tensor
    frr::roi_pool(Tensor input, Tensor rois, float spatial_scale, SymInt pooled_height, SymInt pooled_width) -> (Tensor, Tensor)
    frr::_roi_pool_backward(Tensor grad, Tensor rois, Tensor argmax, float spatial_scale, SymInt pooled_height,
                            SymInt pooled_width, SymInt batch_size, SymInt channels, SymInt height, SymInt width) -> Tensor
    frr::roi_align(Tensor input, Tensor rois, float spatial_scale, SymInt pooled_height, SymInt pooled_width,
                   int sampling_ratio, bool aligned) -> Tensor
    frr::_roi_align_backward(Tensor grad, Tensor rois, float spatial_scale, SymInt pooled_height, SymInt pooled_width,
                             SymInt batch_size, SymInt channels, SymInt height, SymInt width, int sampling_ratio,
                             bool aligned) -> Tensor
    frr::rpn_proposals(Tensor cls, Tensor reg, int image_h, int image_w, int pre_nms_top_k, int post_nms_top_k,
                       float nms_thresh, float min_size) -> (Tensor rois, Tensor count)

Importing this module registers the ops; every implementation is a libfrr kernel (CUDA tensors only, no CPU kernel is
registered: a CPU call fails in the dispatcher).  The fake implementations make the ops traceable (``torch.compile`` /
``make_fx`` / meta tensors) with the right output shapes; ``nms`` has a data-dependent output length (unbacked SymInt).
"""
from __future__ import annotations

import torch

from . import ops, region

_LIB = "frr"


@torch.library.custom_op(f"{_LIB}::nms", mutates_args=(), device_types="cuda")
def nms(dets: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    from .modules import nms as _nms
    return _nms(dets, scores, iou_threshold)


@nms.register_fake
def _(dets, scores, iou_threshold):
    n = torch.library.get_ctx().new_dynamic_size()
    return dets.new_empty((n,), dtype=torch.int64)


@torch.library.custom_op(f"{_LIB}::roi_pool", mutates_args=(), device_types="cuda")
def roi_pool(input: torch.Tensor, rois: torch.Tensor, spatial_scale: float, pooled_height: int,
             pooled_width: int) -> tuple[torch.Tensor, torch.Tensor]:
    feat, _ = ops._feat_layout(input)
    return ops.roi_pool_forward(feat, rois.to(torch.float32).contiguous(), (pooled_height, pooled_width), spatial_scale, True)


@roi_pool.register_fake
def _(input, rois, spatial_scale, pooled_height, pooled_width):
    shape = (rois.shape[0], input.shape[1], pooled_height, pooled_width)
    return input.new_empty(shape), input.new_empty(shape, dtype=torch.int32)


@torch.library.custom_op(f"{_LIB}::_roi_pool_backward", mutates_args=(), device_types="cuda")
def _roi_pool_backward(grad: torch.Tensor, rois: torch.Tensor, argmax: torch.Tensor, spatial_scale: float, pooled_height: int,
                       pooled_width: int, batch_size: int, channels: int, height: int, width: int) -> torch.Tensor:
    return ops.roi_pool_backward(grad.contiguous(), argmax, rois.to(torch.float32).contiguous(),
                                 (batch_size, channels, height, width), spatial_scale, False)


@_roi_pool_backward.register_fake
def _(grad, rois, argmax, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width):
    return grad.new_empty((batch_size, channels, height, width))


def _roi_pool_setup(ctx, inputs, output):
    input, rois, spatial_scale, ph, pw = inputs
    ctx.save_for_backward(rois, output[1])
    ctx.meta = (spatial_scale, ph, pw, tuple(input.shape))
    ctx.mark_non_differentiable(output[1])


def _roi_pool_bwd(ctx, grad_out, _grad_argmax):
    rois, argmax = ctx.saved_tensors
    scale, ph, pw, (b, c, h, w) = ctx.meta
    return torch.ops.frr._roi_pool_backward(grad_out, rois, argmax, scale, ph, pw, b, c, h, w), None, None, None, None


torch.library.register_autograd(f"{_LIB}::roi_pool", _roi_pool_bwd, setup_context=_roi_pool_setup)


@torch.library.custom_op(f"{_LIB}::roi_align", mutates_args=(), device_types="cuda")
def roi_align(input: torch.Tensor, rois: torch.Tensor, spatial_scale: float, pooled_height: int, pooled_width: int,
              sampling_ratio: int, aligned: bool) -> torch.Tensor:
    feat, _ = ops._feat_layout(input)
    return ops.roi_align_forward(feat, rois.to(torch.float32).contiguous(), (pooled_height, pooled_width), spatial_scale,
                                 sampling_ratio, aligned)


@roi_align.register_fake
def _(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio, aligned):
    return input.new_empty((rois.shape[0], input.shape[1], pooled_height, pooled_width))


@torch.library.custom_op(f"{_LIB}::_roi_align_backward", mutates_args=(), device_types="cuda")
def _roi_align_backward(grad: torch.Tensor, rois: torch.Tensor, spatial_scale: float, pooled_height: int, pooled_width: int,
                        batch_size: int, channels: int, height: int, width: int, sampling_ratio: int,
                        aligned: bool) -> torch.Tensor:
    return ops.roi_align_backward(grad.contiguous(), rois.to(torch.float32).contiguous(), (batch_size, channels, height, width),
                                  spatial_scale, sampling_ratio, aligned, False)


@_roi_align_backward.register_fake
def _(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width, sampling_ratio, aligned):
    return grad.new_empty((batch_size, channels, height, width))


def _roi_align_setup(ctx, inputs, output):
    input, rois, spatial_scale, ph, pw, sampling_ratio, aligned = inputs
    ctx.save_for_backward(rois)
    ctx.meta = (spatial_scale, ph, pw, sampling_ratio, aligned, tuple(input.shape))


def _roi_align_bwd(ctx, grad_out):
    (rois,) = ctx.saved_tensors
    scale, ph, pw, sr, aligned, (b, c, h, w) = ctx.meta
    return (torch.ops.frr._roi_align_backward(grad_out, rois, scale, ph, pw, b, c, h, w, sr, aligned),
            None, None, None, None, None, None)


torch.library.register_autograd(f"{_LIB}::roi_align", _roi_align_bwd, setup_context=_roi_align_setup)


@torch.library.custom_op(f"{_LIB}::rpn_proposals", mutates_args=(), device_types="cuda")
def rpn_proposals(cls: torch.Tensor, reg: torch.Tensor, image_h: int, image_w: int, pre_nms_top_k: int, post_nms_top_k: int,
                  nms_thresh: float, min_size: float) -> tuple[torch.Tensor, torch.Tensor]:
    """Batched RegionProposal.forward (models/model.py:17-58): cls [B,N,2] logits or [B,N] scores, reg [B,N,4] ->
    rois [B,post,4] (zero padded), count int32 [B]."""
    rois, count = region.rpn_proposals(cls, reg, image_hw=(image_h, image_w), pre_nms_top_k=pre_nms_top_k,
                                       post_nms_top_k=post_nms_top_k, nms_thresh=nms_thresh, min_size=min_size)
    return rois, count


@rpn_proposals.register_fake
def _(cls, reg, image_h, image_w, pre_nms_top_k, post_nms_top_k, nms_thresh, min_size):
    return reg.new_empty((reg.shape[0], post_nms_top_k, 4)), reg.new_empty((reg.shape[0],), dtype=torch.int32)
