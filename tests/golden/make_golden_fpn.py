"""Golden fixtures for the FPN-variant target makers (run in the BUILD container only).

    python tests/golden/make_golden_fpn.py

Imports the UNMODIFIED ``RPNTargetMaker`` / ``FRCNNTargetMaker`` of /root/reference/models/new_model.py (the
variant main.py trains today) and stores their outputs on seeded synthetic inputs in
tests/golden/targets_fpn.npz.  Same harness-side patches as make_golden.py.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from faster_rcnn_pytorch_b200 import synth  # noqa: E402


def main():
    torch.Tensor.get_device = lambda self: self.device
    sys.modules.setdefault("gdown", types.ModuleType("gdown"))
    sys.path.insert(0, "/root/reference")
    import models.new_model as nm
    import anchor as ref_anchor

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    am = ref_anchor.FRCNNAnchorMaker()
    rtm, ftm = nm.RPNTargetMaker(), nm.FRCNNTargetMaker()
    g = {}
    for name, hw, gseed, G, tseed in [("a", (320, 480), 7000, 5, 7000), ("b", (600, 1000), 7001, 8, 7001),
                                      ("many", (320, 480), 7002, 150, 7002), ("one", (160, 256), 7003, 1, 7003)]:
        anchors = am._enumerate_shifted_anchor(hw)      # any [N,4] anchor set: the maker does not care where it came from
        gt, lab = synth.gt_boxes(gseed, G)
        lab = lab + 1                                    # FPN variant: labels are used as they are (background = 0)
        torch.manual_seed(tseed)
        with contextlib.redirect_stdout(io.StringIO()):
            cls_t, reg_t = rtm(t(gt), t(anchors))
        g[f"{name}_rpn_cls"] = cls_t.numpy().astype(np.int8)
        nz = np.nonzero(cls_t.numpy() >= 0)[0]
        g[f"{name}_rpn_reg_sampled"] = reg_t.numpy()[nz].copy()
        rois_np, _ = synth.random_boxes(tseed + 50, 2000)
        torch.manual_seed(tseed + 1)
        with contextlib.redirect_stdout(io.StringIO()):
            fc, fr, fs = ftm(t(gt), t(lab), t(rois_np))
        g[f"{name}_frcnn_cls"] = fc.numpy().astype(np.int16)
        g[f"{name}_frcnn_reg"] = fr.numpy().copy()
        g[f"{name}_frcnn_rois"] = fs.numpy().copy()
    # a GT that overlaps nothing: the tie-inclusive rule then marks EVERY anchor with IoU == 0 positive (:316-318)
    hw = (160, 256)
    anchors = am._enumerate_shifted_anchor(hw)
    inside = (anchors[:, 0] >= 0) & (anchors[:, 1] >= 0) & (anchors[:, 2] <= 1) & (anchors[:, 3] <= 1)
    anchors_in = anchors[inside]
    gt = np.array([[0.2, 0.2, 0.6, 0.7], [5.0, 5.0, 5.1, 5.1]], np.float32)
    torch.manual_seed(9)
    with contextlib.redirect_stdout(io.StringIO()):
        cls_t, reg_t = rtm(t(gt), t(anchors_in))
    g["far_rpn_cls"] = cls_t.numpy().astype(np.int8)
    g["far_anchors"] = anchors_in.copy()
    np.savez_compressed(os.path.join(HERE, "targets_fpn.npz"), **g)
    print("targets_fpn.npz", os.path.getsize(os.path.join(HERE, "targets_fpn.npz")))


if __name__ == "__main__":
    main()
