"""Generate the golden fixtures in tests/golden/*.npz  (run in the BUILD container only).

    python tests/golden/make_golden.py

Imports the UNMODIFIED reference from /root/reference (models/model.py, anchor.py,
utils/util.py) and torchvision's CPU kernels, runs them on the seeded synthetic inputs of
``faster_rcnn_pytorch_b200.synth`` and stores the outputs.  /root/reference does not exist
on the GPU box, so the fixtures (not the reference) travel with the repo.

Harness-side patches (no edits to the reference):
  1. Tensor.get_device / torch.get_device return the device (the reference's
     ``.to(x.get_device())`` idiom fails on CPU tensors);
  2. stub modules for cv2-independent imports that need network (gdown);
  3. vgg16(pretrained=True) -> vgg16(weights=None).
Full-size cases store seeds + integer outputs + float checksums only (inputs are
regenerated bit-identically from the seed); small cases store complete float outputs.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import os
import sys
import types

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)

from faster_rcnn_pytorch_b200 import synth  # noqa: E402


def import_reference():
    torch.Tensor.get_device = lambda self: self.device
    torch.get_device = lambda t: t.device
    sys.modules.setdefault("gdown", types.ModuleType("gdown"))
    sys.path.insert(0, REF)
    import models.model as ref_model
    import anchor as ref_anchor
    import utils.util as ref_util

    ref_model.vgg16 = lambda pretrained=True: torchvision.models.vgg16(weights=None)
    return ref_model, ref_anchor, ref_util


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def sha(a: np.ndarray) -> np.ndarray:
    h = hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()
    return np.frombuffer(h, dtype=np.uint8).copy()


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def main():
    ref_model, ref_anchor, ref_util = import_reference()
    out = {}
    torch.set_num_threads(8)

    # ------------------------------------------------------------------ anchors (A1/A2)
    am = ref_anchor.FRCNNAnchorMaker()
    g = {"base": am.anchor_base.copy()}
    for hw in [(64, 96), (600, 1000), (608, 1008), (800, 1333), (800, 800)]:
        a = am._enumerate_shifted_anchor(hw)
        key = f"{hw[0]}x{hw[1]}"
        g[f"sha_{key}"] = sha(a)
        g[f"head_{key}"] = a[:27].copy()
        inside = (a[:, 0] >= 0) & (a[:, 1] >= 0) & (a[:, 2] <= 1) & (a[:, 3] <= 1)
        g[f"ninside_{key}"] = np.int64(inside.sum())
        g[f"n_{key}"] = np.int64(a.shape[0])
        if hw == (64, 96):
            g[f"full_{key}"] = a.copy()
    np.savez_compressed(os.path.join(HERE, "anchors.npz"), **g)

    # ------------------------------------------------------------------ box utils (T2/T4/P2)
    rs = np.random.RandomState(11)
    b1, _ = synth.random_boxes(12, 300)
    b2, _ = synth.gt_boxes(13, 7)
    iou = ref_util.find_jaccard_overlap(t(b1), t(b2)).numpy()
    enc = ref_util.encode(ref_util.xy_to_cxcy(t(b2[rs.randint(0, 7, 300)])), ref_util.xy_to_cxcy(t(b1))).numpy()
    treg = (rs.standard_normal((300, 4)) * 0.2).astype(np.float32)
    dec = ref_util.cxcy_to_xy(ref_util.decode(t(treg), ref_util.xy_to_cxcy(t(b1)))).numpy()
    np.savez_compressed(os.path.join(HERE, "boxutils.npz"), iou=iou, enc=enc, dec=dec, treg=treg,
                        gt_pick=rs.get_state()[1][:1])  # inputs regenerated from seeds 11/12/13

    # ------------------------------------------------------------------ NMS (N1/N2) vs torchvision CPU
    g = {}
    kats = {
        "kat1a": (np.array([[0, 0, 5, 1], [2, 0, 10, 1]], np.float32), np.array([0.9, 0.8], np.float32), 0.3),
        "kat1b": (np.array([[0, 0, 7, 1], [0, 0, 10, 1]], np.float32), np.array([0.9, 0.8], np.float32), 0.7),
        "kat2a": (np.array([[0, 0, 1, 1], [2, 2, 3, 3], [4, 4, 5, 5]], np.float32), np.array([0.5, 0.5, 0.5], np.float32), 0.5),
        "kat2b": (np.array([[1, 1, 1, 1], [1, 1, 1, 1]], np.float32), np.array([0.9, 0.8], np.float32), 0.5),
        "kat2c": (np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5),
    }
    for k, (b, s, thr) in kats.items():
        g[k] = torchvision.ops.nms(t(b), t(s), thr).numpy()
    for seed, n, thr in [(100, 64, 0.5), (101, 65, 0.7), (102, 1000, 0.7), (103, 3000, 0.3), (104, 12000, 0.7),
                         (105, 6000, 0.7), (106, 300, 0.3), (107, 1, 0.7), (108, 2500, 0.0), (109, 2500, 1.0)]:
        b, s = synth.random_boxes(seed, n)
        g[f"rand_{seed}_{n}_{thr}"] = torchvision.ops.nms(t(b), t(s), thr).numpy().astype(np.int32)
    # unsorted input with exact score ties (stable order)
    b, s = synth.random_boxes(110, 500)
    s = np.round(s * 8) / 8
    g["ties_110_500_0.5"] = torchvision.ops.nms(t(b), t(s.astype(np.float32)), 0.5).numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **g)

    # ------------------------------------------------------------------ proposal layer (P1-P4,N1)
    rp = ref_model.RegionProposal()
    g = {}
    for name, hw, seed, mode in [("small_train", (160, 256), 200, "train"), ("small_test", (160, 256), 201, "test"),
                                 ("voc_test", (600, 1000), 1000, "test"), ("rpn_train", (608, 1008), 2000, "train"),
                                 ("rpn_train1", (608, 1008), 2001, "train"), ("coco_test", (800, 1333), 4000, "test")]:
        logits, reg, _ = synth.rpn_head_outputs(seed, hw)
        anchor = am._enumerate_shifted_anchor(hw)
        rois = quiet(rp, t(logits), t(reg), t(anchor), mode).numpy()
        g[f"{name}_rois_sha"] = sha(rois)
        g[f"{name}_nrois"] = np.int64(rois.shape[0])
        g[f"{name}_rois_head"] = rois[:64].copy()
        if name.startswith("small"):
            g[f"{name}_rois"] = rois.copy()
        # stage-wise integer outputs from the reference's own intermediate tensors
        sc = torch.softmax(t(logits), dim=-1)[..., 1]
        roi = ref_util.cxcy_to_xy(ref_util.decode(t(reg), ref_util.xy_to_cxcy(t(anchor)))).clamp(0, 1)
        ws, hs = roi[:, 2] - roi[:, 0], roi[:, 3] - roi[:, 1]
        keep = (hs >= (1 / 1000)) & (ws >= (1 / 1000))
        g[f"{name}_nvalid"] = np.int64(keep.sum().item())
        g[f"{name}_valid_packed"] = np.packbits(keep.numpy())
        g[f"{name}_score_sha"] = sha(sc.numpy())
        g[f"{name}_boxes_sha"] = sha(roi.numpy())
        if name.startswith("small"):
            g[f"{name}_score"] = sc.numpy().copy()
            g[f"{name}_boxes"] = roi.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "proposal.npz"), **g)

    # top-k + NMS from IDENTICAL fp32 scores / boxes (tie-free): index-valued goldens
    g = {}
    for name, hw, seed, pre_k, post_k in [("rpn_train", (608, 1008), 2000, 12000, 2000),
                                          ("voc_test", (600, 1000), 1000, 6000, 300),
                                          ("coco_test", (800, 1333), 4000, 6000, 300),
                                          ("small", (160, 256), 200, 12000, 2000)]:
        _, reg, scores = synth.rpn_head_outputs(seed, hw)
        anchor = am._enumerate_shifted_anchor(hw)
        roi = ref_util.cxcy_to_xy(ref_util.decode(t(reg), ref_util.xy_to_cxcy(t(anchor)))).clamp(0, 1)
        ws, hs = roi[:, 2] - roi[:, 0], roi[:, 3] - roi[:, 1]
        keep = (hs >= (1 / 1000)) & (ws >= (1 / 1000))
        roi_c, sc_c = roi[keep, :], t(scores)[keep]
        ss, si = sc_c.sort(descending=True)
        k = min(pre_k, si.numel())
        tb, tsc = roi_c[si[:k]], ss[:k]
        kp = torchvision.ops.nms(tb, tsc, 0.7)[:post_k]
        g[f"{name}_topk_idx"] = si[:k].numpy().astype(np.int32)
        g[f"{name}_keep"] = kp.numpy().astype(np.int32)
        g[f"{name}_boxes_sha"] = sha(roi.numpy())
        g[f"{name}_rois_sha"] = sha(tb[kp].numpy())
    np.savez_compressed(os.path.join(HERE, "topk_nms.npz"), **g)

    # ------------------------------------------------------------------ target makers (T1, T3)
    rtm, ftm = ref_model.RPNTargetMaker(), ref_model.FastRcnnTargetMaker()
    g = {}
    cases = [("kat6", (600, 1000), None, 2, 7), ("c3_0", (600, 1000), 3000, 8, 3000), ("c3_1", (600, 1000), 3001, 8, 3001),
             ("many_gt", (600, 1000), 3100, 160, 3100), ("one_gt", (320, 480), 3200, 1, 3200),
             ("small", (160, 256), 3300, 3, 3300)]
    for name, hw, gseed, G, tseed in cases:
        anchor = am._enumerate_shifted_anchor(hw)
        if gseed is None:
            gt = np.array([[0.1, 0.2, 0.5, 0.7], [0.3, 0.3, 0.9, 0.95]], np.float32)
            lab = np.array([11, 14], np.int64)
        else:
            gt, lab = synth.gt_boxes(gseed, G)
        torch.manual_seed(tseed)
        cls_t, reg_t = quiet(rtm, t(gt), t(anchor))
        g[f"{name}_rpn_cls"] = cls_t.numpy().astype(np.int8)
        g[f"{name}_rpn_reg_sha"] = sha(reg_t.numpy())
        if name in ("small", "one_gt"):
            g[f"{name}_rpn_reg"] = reg_t.numpy().copy()
        else:
            nz = np.nonzero(cls_t.numpy() >= 0)[0]
            g[f"{name}_rpn_reg_sampled"] = reg_t.numpy()[nz].copy()
        # proposals for the Fast R-CNN target maker: 2000 seeded boxes (no exp(): bit-stable inputs)
        rois_np, _ = synth.random_boxes(tseed + 50, 2000)
        rois = t(rois_np)
        torch.manual_seed(tseed + 1)
        fc, fr, fs = quiet(ftm, [t(gt)], [t(lab)], rois)
        g[f"{name}_frcnn_cls"] = fc.numpy().astype(np.int16)
        g[f"{name}_frcnn_reg"] = fr.numpy().copy()
        g[f"{name}_frcnn_rois"] = fs.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "targets.npz"), **g)

    # ------------------------------------------------------------------ RoIPool / RoIAlign (R1-R4)
    g = {}
    for name, seed, B, C, fh, fw, K in [("a", 500, 1, 8, 37, 62, 64), ("b", 501, 2, 16, 10, 14, 40),
                                        ("c", 502, 1, 4, 5, 6, 30)]:
        feat = synth.features(seed, B, C, fh, fw)
        rois5 = synth.random_rois(seed + 1, K, fh, fw, B)
        go = np.random.RandomState(seed + 2).standard_normal((K, C, 7, 7)).astype(np.float32)
        o, am_ = torch.ops.torchvision.roi_pool(t(feat), t(rois5), 1.0, 7, 7)
        gi = torch.ops.torchvision._roi_pool_backward(t(go), t(rois5), am_, 1.0, 7, 7, B, C, fh, fw)
        g[f"{name}_pool_out"] = o.numpy().copy()
        g[f"{name}_pool_argmax"] = am_.numpy().astype(np.int32)
        g[f"{name}_pool_gin"] = gi.numpy().copy()
        for scale in (1.0, 0.5):
            oa = torch.ops.torchvision.roi_align(t(feat), t(rois5), scale, 7, 7, 2, False)
            gia = torch.ops.torchvision._roi_align_backward(t(go), t(rois5), scale, 7, 7, B, C, fh, fw, 2, False)
            g[f"{name}_align_out_{scale}"] = oa.numpy().copy()
            g[f"{name}_align_gin_{scale}"] = gia.numpy().copy()
    # the reference's own call path: FastRCNNHead scaling + RoIPool module on a list of rois
    feat = synth.features(510, 1, 8, 37, 62)
    b, _ = synth.random_boxes(511, 32)
    pool = torchvision.ops.RoIPool(output_size=(7, 7), spatial_scale=1.0)
    scaled = t(b) * torch.FloatTensor([62, 37, 62, 37])
    g["head_pool_out"] = pool(t(feat), [scaled]).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "roi.npz"), **g)

    # ------------------------------------------------------------------ predict tail (D1, D2)
    g = {}
    model = quiet(ref_model.FRCNN, 21)
    model.eval()
    torch.manual_seed(77)
    with torch.no_grad():
        model.fast_rcnn_head.cls_head.weight.normal_(0, 0.03)
        model.fast_rcnn_head.reg_head.weight.normal_(0, 0.01)
    cap = {}
    model.fast_rcnn_head.register_forward_hook(lambda m, i, o: cap.update(cls=o[0].detach(), reg=o[1].detach(), rois=i[1].detach()))
    x = torch.from_numpy(np.random.RandomState(78).standard_normal((1, 3, 224, 320)).astype(np.float32))
    for thres in (0.05, 0.005):
        with torch.no_grad():
            bbox, label, score = quiet(model.predict, x, types.SimpleNamespace(thres=thres))
        g[f"det_bbox_{thres}"] = bbox
        g[f"det_label_{thres}"] = label
        g[f"det_score_{thres}"] = score
    g["head_cls"] = cap["cls"].numpy().copy()
    g["head_reg"] = cap["reg"].numpy().copy()
    g["head_rois"] = cap["rois"].numpy().copy()
    # _suppress alone on synthetic, well-spread inputs (C=21 and C=81)
    for C, seed in [(21, 600), (81, 601)]:
        cls, reg = synth.head_outputs(seed, 300, C)
        rois, _ = synth.random_boxes(seed + 1, 300)
        prob = torch.softmax(t(cls), dim=-1)
        r = t(reg).reshape(-1, C, 4) * torch.FloatTensor([0.1, 0.1, 0.2, 0.2])
        rr = t(rois).reshape(-1, 1, 4).expand_as(r)
        pb = ref_util.cxcy_to_xy(ref_util.decode(r.reshape(-1, 4), ref_util.xy_to_cxcy(rr.reshape(-1, 4))))
        pb = pb.reshape(-1, C * 4).clamp(min=0, max=1)
        g[f"sup{C}_prob"] = prob.numpy().copy()
        g[f"sup{C}_boxes_head"] = pb.numpy()[:8].copy()
        # _suppress itself is pinned on exp()-free boxes so both sides see identical fp32 inputs
        pb = t(synth.random_boxes(seed + 2, 300 * C, cluster=False)[0].reshape(300, C * 4))
        for thres in (0.05, 0.005):
            fake_self = types.SimpleNamespace(num_classes=C)
            bb, ll, ss = ref_model.FRCNN._suppress(fake_self, pb, prob, types.SimpleNamespace(thres=thres))
            g[f"sup{C}_bbox_{thres}"] = bb
            g[f"sup{C}_label_{thres}"] = ll
            g[f"sup{C}_score_{thres}"] = ss
    np.savez_compressed(os.path.join(HERE, "predict.npz"), **g)

    # ------------------------------------------------------------------ host RNG (KAT-5)
    g = {}
    for seed, ns in [(123, (10,)), (5, (5000, 7, 0, 1, 2, 300))]:
        torch.manual_seed(seed)
        for i, n in enumerate(ns):
            g[f"perm_{seed}_{i}_{n}"] = torch.randperm(n).numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "rng.npz"), **g)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
