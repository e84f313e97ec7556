"""GPU parity: predict tail (D1 decode_classwise) and FRCNN._suppress (D2 class_nms) through the C ABI
vs the reference goldens and the CPU oracle.  Labels / kept sets / scores bit-exact from identical fp32
inputs; softmax / exp()-bearing boxes within 1e-5."""
import numpy as np
import pytest
import torch

from conftest import golden
from faster_rcnn_pytorch_b200 import modules, ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def close(a, b, atol=1e-6):
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=atol)


@pytest.mark.parametrize("C,seed", [(21, 600), (81, 601)])
def test_decode_classwise_vs_reference(oracle, C, seed):
    g = golden("predict")
    cls, reg = synth.head_outputs(seed, 300, C)
    rois, _ = synth.random_boxes(seed + 1, 300)
    prob, boxes = ops.decode_classwise(dev(cls), dev(reg), dev(rois), C)
    close(prob.cpu().numpy(), g[f"sup{C}_prob"], atol=1e-7)
    close(boxes.cpu().numpy()[:8], g[f"sup{C}_boxes_head"])
    wp, wb = oracle.decode_classwise(cls, reg, rois, C)
    close(prob.cpu().numpy(), wp, atol=1e-7)
    close(boxes.cpu().numpy(), wb)


@pytest.mark.parametrize("C,seed", [(21, 600), (81, 601)])
@pytest.mark.parametrize("thres", [0.05, 0.005])
def test_suppress_bit_exact_vs_reference(oracle, C, seed, thres):
    """Identical fp32 prob + boxes on both sides -> class-major detections identical to FRCNN._suppress."""
    g = golden("predict")
    boxes = synth.random_boxes(seed + 2, 300 * C, cluster=False)[0].reshape(300, C * 4)
    prob = g[f"sup{C}_prob"]
    bb, ll, ss = modules.suppress(dev(boxes), dev(prob), C, thres)
    assert ll.dtype == np.int32 and bb.dtype == np.float32 and ss.dtype == np.float32
    assert np.array_equal(ll, g[f"sup{C}_label_{thres}"])
    assert np.array_equal(ss, g[f"sup{C}_score_{thres}"])
    assert np.array_equal(bb, g[f"sup{C}_bbox_{thres}"])


@pytest.mark.parametrize("thres", [0.05, 0.005])
def test_predict_tail_from_reference_head_outputs(oracle, thres):
    """Head outputs captured from the reference's FRCNN.predict: decode on the GPU, then _suppress.  The
    kept set is index-valued, so it is checked against the oracle fed with the GPU's own fp32 prob/boxes."""
    g = golden("predict")
    prob, boxes = modules.predict_tail(dev(g["head_cls"]), dev(g["head_reg"]), dev(g["head_rois"]), 21)
    wp, wb = oracle.decode_classwise(g["head_cls"], g["head_reg"], g["head_rois"], 21)
    close(prob.cpu().numpy(), wp, atol=1e-7)
    close(boxes.cpu().numpy(), wb)
    bb, ll, ss = modules.suppress(boxes, prob, 21, thres)
    ob, ol, os_ = oracle.suppress(boxes.cpu().numpy(), prob.cpu().numpy(), 21, thres)
    assert np.array_equal(ll, ol) and np.array_equal(ss, os_) and np.array_equal(bb, ob)
    gl = g[f"det_label_{thres}"]
    if len(gl) == len(ll):      # the reference's own detections (exp/softmax ulps may flip a borderline score)
        assert np.array_equal(ll, gl)
        close(ss, g[f"det_score_{thres}"])
        close(bb, g[f"det_bbox_{thres}"])


def test_class_nms_batched_counts_caps_and_ties(oracle):
    B, R, C = 3, 300, 21
    rs = np.random.RandomState(900)
    prob = np.stack([oracle.softmax_rows(rs.standard_normal((R, C)).astype(np.float32) * 2) for _ in range(B)])
    prob = (np.round(prob * 64) / 64).astype(np.float32)          # many exact score ties -> stable order matters
    boxes = np.stack([synth.random_boxes(910 + i, R * C, cluster=True)[0].reshape(R, C * 4) for i in range(B)])
    rc = np.asarray([300, 120, 0], np.int32)
    db, dl, ds, dc = ops.class_nms(dev(prob), dev(boxes), C, score_thres=0.05, iou_thr=0.3, roi_count=dev(rc))
    dc = dc.cpu().numpy()
    for i in range(B):
        ob, ol, os_ = oracle.suppress(boxes[i, :rc[i]], prob[i, :rc[i]], C, 0.05)
        assert dc[i] == len(ol)
        assert np.array_equal(dl[i, :dc[i]].cpu().numpy(), ol)
        assert np.array_equal(ds[i, :dc[i]].cpu().numpy(), os_)
        assert np.array_equal(db[i, :dc[i]].cpu().numpy(), ob)
        assert (dl[i, dc[i]:].cpu().numpy() == -1).all()
    # capacity: truncates class-major, count is clamped
    cap = 37
    db2, dl2, ds2, dc2 = ops.class_nms(dev(prob), dev(boxes), C, score_thres=0.05, iou_thr=0.3, roi_count=dev(rc), cap=cap)
    assert dc2.cpu().numpy().tolist() == [min(int(x), cap) for x in dc]
    assert torch.equal(dl2[0], dl[0, :cap]) and torch.equal(db2[0], db[0, :cap])


def test_class_nms_iou_threshold_is_compared_in_double(oracle):
    """KAT-1 through the per-class path: IoU == fl32(0.3) IS suppressed at iou_thr 0.3 (0.3f > 0.3)."""
    C = 2
    boxes = np.zeros((2, C, 4), np.float32)
    boxes[:, 1] = [[0, 0, 5, 1], [2, 0, 10, 1]]
    prob = np.array([[0.1, 0.9], [0.2, 0.8]], np.float32)
    bb, ll, ss = modules.suppress(dev(boxes.reshape(2, C * 4)), dev(prob), C, 0.05)
    assert ll.tolist() == [0] and ss.tolist() == [np.float32(0.9)]


def test_pack_detections_kernel_scaling_xywh_and_truncation():
    """Evaluation hand-off: pixel scaling (test.py:68-71), xyxy -> xywh (coco_eval.py:156-158), fixed-shape rows."""
    from faster_rcnn_pytorch_b200 import dist as fdist
    rs = np.random.RandomState(5)
    B, cap = 4, 30
    counts = np.asarray([30, 12, 0, 7], np.int32)
    boxes = rs.uniform(0, 1, (B, cap, 4)).astype(np.float32)
    labels = rs.randint(0, 80, (B, cap)).astype(np.int32)
    scores = rs.uniform(0.05, 1, (B, cap)).astype(np.float32)
    wh = np.asarray([[1333, 800], [640, 480], [500, 375], [1000, 600]], np.float32)
    for max_det, xywh, use_wh in [(100, True, True), (10, False, True), (10, True, False)]:
        out, cnt = fdist.pack_detections(dev(boxes), dev(labels), dev(scores), dev(counts), max_det,
                                         image_wh=dev(wh) if use_wh else None, xywh=xywh)
        out, cnt = out.cpu().numpy(), cnt.cpu().numpy()
        for i in range(B):
            k = min(int(counts[i]), max_det, cap)
            assert cnt[i] == k
            sc = np.asarray([wh[i, 0], wh[i, 1], wh[i, 0], wh[i, 1]], np.float32) if use_wh else np.ones(4, np.float32)
            b = boxes[i, :k] * sc
            if xywh:
                b = np.stack([b[:, 0], b[:, 1], b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], axis=1)
            assert np.array_equal(out[i, :k, :4], b.astype(np.float32))
            assert np.array_equal(out[i, :k, 4], scores[i, :k]) and np.array_equal(out[i, :k, 5], labels[i, :k].astype(np.float32))
            assert (out[i, k:] == 0).all()
