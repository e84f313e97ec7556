"""Reference-shaped VGG16 Faster R-CNN training step around the libfrr region stage (BASELINE configs[4]).

NOT part of the product: the backbone, the RPN convs and the FC head are plain torch / cuDNN / cuBLAS modules with the
structure of the reference's ``FRCNN`` (models/model.py:60-121,268-343); everything between the stride-16 feature map and
the losses goes through ``faster_rcnn_pytorch_b200`` (batched, no host synchronisation):

    features -> RPN convs (channels_last, zero-copy [B,N,2]/[B,N,4] views) -> region.rpn_proposals (decode + top-k + NMS)
             -> targets.make_targets (device-side sampling) -> ops.roi_pool (autograd) -> FC head -> ops.region_loss

Used by ``bench.py --workload joint`` (torch DDP over NCCL for the backbone gradients) and by
``tests/test_gpu_modules.py::test_joint_training_step``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torchvision

from faster_rcnn_pytorch_b200 import ops, region, targets


class RegionTimer:
    """CUDA-event brackets around the region-stage calls of one step (summed after a synchronise)."""

    def __init__(self, enabled: bool):
        self.enabled, self.pairs = enabled, []

    def __call__(self, fn, *a, **kw):
        if not self.enabled:
            return fn(*a, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **kw)
        e1.record()
        self.pairs.append((e0, e1))
        return out

    def total_ms(self):
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self.pairs)
        self.pairs = []
        return ms


class FRCNNTrain(nn.Module):
    def __init__(self, num_classes: int = 21, width_div: int = 1):
        super().__init__()
        vgg = torchvision.models.vgg16(weights=None)                     # random init: no network for checkpoints
        self.extractor = nn.Sequential(*list(vgg.features.children())[:-1])          # models/model.py:279-281
        self.classifier = nn.Sequential(nn.Linear(25088, 4096 // width_div), nn.ReLU(inplace=True),
                                        nn.Linear(4096 // width_div, 4096 // width_div), nn.ReLU(inplace=True))
        self.inter_layer = nn.Conv2d(512, 512, 3, padding=1)                          # models/model.py:67-69
        self.cls_layer = nn.Conv2d(512, 18, 1)
        self.reg_layer = nn.Conv2d(512, 36, 1)
        self.cls_head = nn.Linear(4096 // width_div, num_classes)                     # models/model.py:93-94
        self.reg_head = nn.Linear(4096 // width_div, num_classes * 4)
        for m, std in ((self.inter_layer, 0.01), (self.cls_layer, 0.01), (self.reg_layer, 0.01), (self.cls_head, 0.01),
                       (self.reg_head, 0.001)):
            nn.init.normal_(m.weight, 0, std)
            nn.init.zeros_(m.bias)
        self.num_classes = num_classes
        self.timer = RegionTimer(False)

    def forward(self, x, gt, gt_label, generator):
        """x [B,3,H,W], gt [B,G,4] normalised xyxy, gt_label int64 [B,G] -> loss [B,5] (losses/loss.py:62-82 per image)."""
        t = self.timer
        B, _, H, W = x.shape
        feats = self.extractor(x)                                                     # [B,512,H/16,W/16]
        h = torch.relu(self.inter_layer(feats))
        cls, reg = region.rpn_head_views(self.cls_layer(h), self.reg_layer(h))        # zero copy under channels_last
        with torch.no_grad():
            rois, count = t(region.rpn_proposals, cls.detach(), reg.detach(), image_hw=(H, W), mode="train")
            tg = t(targets.make_targets, gt, None, gt_label, rois, count, image_hw=(H, W), generator=generator)
            fh, fw = feats.shape[2], feats.shape[3]
            scale = torch.tensor([fw, fh, fw, fh], dtype=torch.float32, device=x.device)
            S = tg["sample_rois"].shape[1]
            bidx = torch.arange(B, device=x.device, dtype=torch.float32).repeat_interleave(S)[:, None]
            rois5 = torch.cat([bidx, (tg["sample_rois"] * scale).reshape(-1, 4)], dim=1)
        pool = t(ops.roi_pool, feats, rois5, (7, 7), 1.0)                             # [B*S,512,7,7], autograd
        y = self.classifier(pool.flatten(1))
        head_cls = self.cls_head(y).view(B, S, self.num_classes)
        head_reg = self.reg_head(y).view(B, S, self.num_classes, 4)
        return t(ops.region_loss, cls, reg, tg["rpn_cls"], tg["rpn_reg"], head_cls, head_reg, tg["frcnn_cls"],
                 tg["frcnn_reg"])


def make_batch(B, hw, G, device, seed=0):
    from faster_rcnn_pytorch_b200 import synth
    import numpy as np
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.standard_normal((B, 3, hw[0], hw[1])).astype(np.float32)).to(device)
    gt = torch.from_numpy(np.stack([synth.gt_boxes(seed * 100 + i, G)[0] for i in range(B)])).to(device)
    lab = torch.from_numpy(np.stack([synth.gt_boxes(seed * 100 + i, G)[1] for i in range(B)])).to(device)
    return x.contiguous(memory_format=torch.channels_last), gt, lab + 1 - 1
