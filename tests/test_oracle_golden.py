"""Pins the CPU oracle (oracle/) against the golden vectors generated from the real reference
+ torchvision CPU kernels (tests/golden/make_golden.py) and the KATs of SURVEY.md §8c."""
import numpy as np
import pytest

from conftest import golden, sha
from faster_rcnn_pytorch_b200 import synth

RTOL = 1e-5  # north_star tolerance for exp/log-bearing outputs


def close(a, b, rtol=RTOL, atol=1e-6):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


# ------------------------------------------------------------------------------ anchors
def test_anchor_base_kat3(oracle):
    g = golden("anchors")
    tab = oracle.anchor_base_table()
    assert np.array_equal(tab, g["base"])
    np.testing.assert_allclose(tab[0], [-37.254833, -82.50967, 53.254833, 98.50967], rtol=1e-6)
    assert np.array_equal(tab[3], np.array([-56, -56, 72, 72], np.float32))


@pytest.mark.parametrize("hw,ninside", [((64, 96), None), ((600, 1000), 8044), ((608, 1008), 8138),
                                        ((800, 1333), 18082), ((800, 800), 8940)])
def test_anchor_enumeration_bit_exact(oracle, hw, ninside):
    g = golden("anchors")
    key = f"{hw[0]}x{hw[1]}"
    a = oracle.enumerate_anchors(hw)
    assert a.shape[0] == int(g[f"n_{key}"]) == synth.num_anchors(hw)
    assert np.array_equal(sha(a), g[f"sha_{key}"])
    assert np.array_equal(a[:27], g[f"head_{key}"])
    inside = (a[:, 0] >= 0) & (a[:, 1] >= 0) & (a[:, 2] <= 1) & (a[:, 3] <= 1)
    assert int(inside.sum()) == int(g[f"ninside_{key}"])
    if ninside is not None:
        assert int(inside.sum()) == ninside


# ------------------------------------------------------------------------------ box utils
def test_box_utils(oracle):
    g = golden("boxutils")
    rs = np.random.RandomState(11)
    b1, _ = synth.random_boxes(12, 300)
    b2, _ = synth.gt_boxes(13, 7)
    assert np.array_equal(oracle.pairwise_iou_eps(b1, b2), g["iou"])          # bit-exact
    pick = rs.randint(0, 7, 300)
    enc = oracle.encode_boxes(oracle.corners_to_center(b2[pick]), oracle.corners_to_center(b1))
    close(enc, g["enc"])
    treg = (rs.standard_normal((300, 4)) * 0.2).astype(np.float32)
    assert np.array_equal(treg, g["treg"])
    close(oracle.center_to_corners(oracle.decode_boxes(treg, oracle.corners_to_center(b1))), g["dec"])


# ------------------------------------------------------------------------------ NMS
def test_nms_kats(oracle):
    g = golden("nms")
    f = np.float32
    assert oracle.nms(np.array([[0, 0, 5, 1], [2, 0, 10, 1]], f), np.array([.9, .8], f), 0.3).tolist() == [0]
    assert oracle.nms(np.array([[0, 0, 7, 1], [0, 0, 10, 1]], f), np.array([.9, .8], f), 0.7).tolist() == [0, 1]
    assert oracle.nms(np.array([[0, 0, 1, 1], [2, 2, 3, 3], [4, 4, 5, 5]], f), np.array([.5, .5, .5], f), 0.5).tolist() == [0, 1, 2]
    assert oracle.nms(np.array([[1, 1, 1, 1], [1, 1, 1, 1]], f), np.array([.9, .8], f), 0.5).tolist() == [0, 1]
    e = oracle.nms(np.zeros((0, 4), f), np.zeros((0,), f), 0.5)
    assert e.dtype == np.int64 and e.shape == (0,)
    for k in ("kat1a", "kat1b", "kat2a", "kat2b", "kat2c"):
        assert g[k].dtype == np.int64


@pytest.mark.parametrize("key", ["rand_100_64_0.5", "rand_101_65_0.7", "rand_102_1000_0.7", "rand_103_3000_0.3",
                                 "rand_104_12000_0.7", "rand_105_6000_0.7", "rand_106_300_0.3", "rand_107_1_0.7",
                                 "rand_108_2500_0.0", "rand_109_2500_1.0"])
def test_nms_random_vs_torchvision_golden(oracle, key):
    g = golden("nms")
    _, seed, n, thr = key.split("_")
    b, s = synth.random_boxes(int(seed), int(n))
    k = oracle.nms(b, s, float(thr))
    assert np.array_equal(k, g[key].astype(np.int64))
    if int(n) <= 1000:   # the numpy fallback states the same algorithm
        order = np.argsort(-s.astype(np.float64), kind="stable")
        assert np.array_equal(oracle._nms_numpy(b, order, float(thr)), k)


def test_nms_score_ties_are_stable(oracle):
    g = golden("nms")
    b, s = synth.random_boxes(110, 500)
    s = (np.round(s * 8) / 8).astype(np.float32)
    assert np.array_equal(oracle.nms(b, s, 0.5), g["ties_110_500_0.5"].astype(np.int64))


# ------------------------------------------------------------------------------ proposal layer
@pytest.mark.parametrize("name,hw,seed,mode", [("small_train", (160, 256), 200, "train"),
                                               ("small_test", (160, 256), 201, "test")])
def test_region_proposal_small_full(oracle, name, hw, seed, mode):
    g = golden("proposal")
    logits, reg, _ = synth.rpn_head_outputs(seed, hw)
    r = oracle.region_proposal(logits, reg, oracle.enumerate_anchors(hw), mode)
    close(r["score"], g[f"{name}_score"])
    close(r["boxes"], g[f"{name}_boxes"])
    assert int(r["valid"].sum()) == int(g[f"{name}_nvalid"])
    assert np.array_equal(np.packbits(r["valid"]), g[f"{name}_valid_packed"])
    assert r["rois"].shape[0] == int(g[f"{name}_nrois"])
    close(r["rois"], g[f"{name}_rois"])


@pytest.mark.parametrize("name,hw,seed,pre_k,post_k", [("rpn_train", (608, 1008), 2000, 12000, 2000),
                                                       ("voc_test", (600, 1000), 1000, 6000, 300),
                                                       ("coco_test", (800, 1333), 4000, 6000, 300),
                                                       ("small", (160, 256), 200, 12000, 2000)])
def test_topk_and_nms_indices_bit_exact(oracle, name, hw, seed, pre_k, post_k):
    """Index-valued outputs from identical fp32 scores: top-k indices (into the compacted array,
    as the reference) and NMS keep positions must be bit-exact."""
    g = golden("topk_nms")
    _, reg, scores = synth.rpn_head_outputs(seed, hw)
    mode = "train" if pre_k == 12000 else "test"
    r = oracle.region_proposal(None, reg, oracle.enumerate_anchors(hw), mode, scores=scores)
    assert np.array_equal(r["topk_idx"], g[f"{name}_topk_idx"].astype(np.int64))
    assert np.array_equal(r["keep"], g[f"{name}_keep"].astype(np.int64))


# ------------------------------------------------------------------------------ host RNG
def test_host_randperm_matches_torch_stream(oracle):
    g = golden("rng")
    rp = oracle.HostRandperm(123)
    assert rp(10).tolist() == [2, 0, 8, 1, 3, 7, 4, 9, 5, 6] == g["perm_123_0_10"].tolist()
    rp = oracle.HostRandperm(5)
    for i, n in enumerate((5000, 7, 0, 1, 2, 300)):
        assert np.array_equal(rp(n), g[f"perm_5_{i}_{n}"].astype(np.int64))


# ------------------------------------------------------------------------------ target makers
CASES = [("kat6", (600, 1000), None, 2, 7), ("c3_0", (600, 1000), 3000, 8, 3000), ("c3_1", (600, 1000), 3001, 8, 3001),
         ("many_gt", (600, 1000), 3100, 160, 3100), ("one_gt", (320, 480), 3200, 1, 3200), ("small", (160, 256), 3300, 3, 3300)]


def _gt(gseed, G):
    if gseed is None:
        return np.array([[0.1, 0.2, 0.5, 0.7], [0.3, 0.3, 0.9, 0.95]], np.float32), np.array([11, 14], np.int64)
    return synth.gt_boxes(gseed, G)


@pytest.mark.parametrize("name,hw,gseed,G,tseed", CASES)
def test_rpn_targets(oracle, name, hw, gseed, G, tseed):
    g = golden("targets")
    gt, _ = _gt(gseed, G)
    r = oracle.rpn_targets(gt, oracle.enumerate_anchors(hw), oracle.HostRandperm(tseed))
    assert np.array_equal(r["labels"], g[f"{name}_rpn_cls"].astype(np.int64))      # bit-exact labels
    if f"{name}_rpn_reg" in g:
        close(r["reg"], g[f"{name}_rpn_reg"])
    else:
        close(r["reg"][r["labels"] >= 0], g[f"{name}_rpn_reg_sampled"])
    if name == "kat6":
        lab = r["labels"]
        assert ((lab == 1).sum(), (lab == 0).sum(), (lab == -1).sum()) == (41, 215, 20390)
    if name == "many_gt":
        assert r["n_pos"] > 128 and (r["labels"] == 1).sum() == 128 and (r["labels"] == 0).sum() == 128


def test_rpn_targets_rejects_empty_gt(oracle):
    with pytest.raises(IndexError):
        oracle.rpn_targets(np.zeros((0, 4), np.float32), oracle.enumerate_anchors((160, 256)), oracle.HostRandperm(0))


@pytest.mark.parametrize("name,hw,gseed,G,tseed", CASES)
def test_frcnn_targets(oracle, name, hw, gseed, G, tseed):
    g = golden("targets")
    gt, lab = _gt(gseed, G)
    rois, _ = synth.random_boxes(tseed + 50, 2000)
    r = oracle.frcnn_targets(gt, lab, rois, oracle.HostRandperm(tseed + 1))
    assert np.array_equal(r["cls"], g[f"{name}_frcnn_cls"].astype(np.int64))      # sampled indices bit-exact
    assert np.array_equal(r["sample_rois"], g[f"{name}_frcnn_rois"])
    close(r["reg"], g[f"{name}_frcnn_reg"], atol=2e-5)
    assert r["cls"].shape == (128,) and r["reg"].shape == (128, 4)


# ------------------------------------------------------------------------------ RoIPool / RoIAlign
@pytest.mark.parametrize("name,seed,B,C,fh,fw,K", [("a", 500, 1, 8, 37, 62, 64), ("b", 501, 2, 16, 10, 14, 40),
                                                   ("c", 502, 1, 4, 5, 6, 30)])
def test_roi_pool_align_bit_exact(oracle, name, seed, B, C, fh, fw, K):
    g = golden("roi")
    feat = synth.features(seed, B, C, fh, fw)
    rois5 = synth.random_rois(seed + 1, K, fh, fw, B)
    go = np.random.RandomState(seed + 2).standard_normal((K, C, 7, 7)).astype(np.float32)
    out, arg = oracle.roi_pool_forward(feat, rois5)
    assert np.array_equal(out, g[f"{name}_pool_out"])
    assert np.array_equal(arg, g[f"{name}_pool_argmax"])
    assert np.array_equal(oracle.roi_pool_backward(go, arg, rois5, feat.shape), g[f"{name}_pool_gin"])
    for scale in (1.0, 0.5):
        assert np.array_equal(oracle.roi_align_forward(feat, rois5, spatial_scale=scale), g[f"{name}_align_out_{scale}"])
        close(oracle.roi_align_backward(go, rois5, feat.shape, spatial_scale=scale), g[f"{name}_align_gin_{scale}"])


def test_head_roi_pool_call_path(oracle):
    g = golden("roi")
    feat = synth.features(510, 1, 8, 37, 62)
    b, _ = synth.random_boxes(511, 32)
    out, _ = oracle.roi_pool_forward(feat, oracle.scale_rois(b, 37, 62))
    assert np.array_equal(out, g["head_pool_out"])


# ------------------------------------------------------------------------------ predict tail
@pytest.mark.parametrize("thres", [0.05, 0.005])
def test_predict_tail_from_reference_head_outputs(oracle, thres):
    g = golden("predict")
    prob, boxes = oracle.decode_classwise(g["head_cls"], g["head_reg"], g["head_rois"], 21)
    bb, ll, ss = oracle.suppress(boxes, prob, 21, thres)
    gb, gl, gs = g[f"det_bbox_{thres}"], g[f"det_label_{thres}"], g[f"det_score_{thres}"]
    assert gl.dtype == np.int32 and gb.dtype == np.float32
    assert np.array_equal(ll, gl)
    close(ss, gs)
    close(bb, gb)


@pytest.mark.parametrize("C,seed", [(21, 600), (81, 601)])
@pytest.mark.parametrize("thres", [0.05, 0.005])
def test_suppress_bit_exact_from_identical_inputs(oracle, C, seed, thres):
    g = golden("predict")
    cls, reg = synth.head_outputs(seed, 300, C)
    rois, _ = synth.random_boxes(seed + 1, 300)
    prob_o, boxes = oracle.decode_classwise(cls, reg, rois, C)
    close(prob_o, g[f"sup{C}_prob"])
    close(boxes[:8], g[f"sup{C}_boxes_head"])
    # _suppress: the reference's own fp32 prob + exp()-free seeded boxes on both sides -> bit-exact
    boxes = synth.random_boxes(seed + 2, 300 * C, cluster=False)[0].reshape(300, C * 4)
    bb, ll, ss = oracle.suppress(boxes, g[f"sup{C}_prob"], C, thres)
    assert np.array_equal(ll, g[f"sup{C}_label_{thres}"])
    assert np.array_equal(ss, g[f"sup{C}_score_{thres}"])
    assert np.array_equal(bb, g[f"sup{C}_bbox_{thres}"])
    assert len(ll) > 0


# ------------------------------------------------------------------------------ FPN-variant target makers
FPN_CASES = [("a", (320, 480), 7000, 5, 7000), ("b", (600, 1000), 7001, 8, 7001), ("many", (320, 480), 7002, 150, 7002),
             ("one", (160, 256), 7003, 1, 7003)]


@pytest.mark.parametrize("name,hw,gseed,G,tseed", FPN_CASES)
def test_fpn_variant_targets(oracle, name, hw, gseed, G, tseed):
    """models/new_model.py:153-206,299-349: eps-free box_iou, no inside filter, tie-inclusive match, 512/128 sampling."""
    g = golden("targets_fpn")
    anchors = oracle.enumerate_anchors(hw)
    gt, lab = synth.gt_boxes(gseed, G)
    r = oracle.rpn_targets(gt, anchors, oracle.HostRandperm(tseed), variant="fpn")
    assert np.array_equal(r["labels"], g[f"{name}_rpn_cls"].astype(np.int64))
    close(r["reg"][r["labels"] >= 0], g[f"{name}_rpn_reg_sampled"])
    rois, _ = synth.random_boxes(tseed + 50, 2000)
    f = oracle.frcnn_targets(gt, lab + 1, rois, oracle.HostRandperm(tseed + 1), variant="fpn")
    assert f["cls"].shape == (512,)
    assert np.array_equal(f["cls"], g[f"{name}_frcnn_cls"].astype(np.int64))
    assert np.array_equal(f["sample_rois"], g[f"{name}_frcnn_rois"])
    close(f["reg"], g[f"{name}_frcnn_reg"], atol=2e-5)


def test_fpn_variant_tie_inclusive_zero_iou(oracle):
    g = golden("targets_fpn")
    gt = np.array([[0.2, 0.2, 0.6, 0.7], [5.0, 5.0, 5.1, 5.1]], np.float32)
    r = oracle.rpn_targets(gt, g["far_anchors"], oracle.HostRandperm(9), variant="fpn")
    assert np.array_equal(r["labels"], g["far_rpn_cls"].astype(np.int64))


@pytest.mark.parametrize("name,mode", [("train", "train"), ("test", "test")])
def test_fpn_variant_proposals(oracle, name, mode):
    """models/new_model.py:46-83: min_size 10/1000, 4000|2000 -> 1000, torchvision multi-level anchors."""
    g = golden("proposal_fpn")
    pre_k, post_k = (4000, 1000) if mode == "train" else (2000, 1000)
    boxes = oracle.decode_clip(g[f"{name}_reg"], g[f"{name}_anchor"])
    close(boxes, g[f"{name}_boxes"])
    valid = oracle.min_size_mask(g[f"{name}_boxes"], 10.0)
    assert np.array_equal(valid, g[f"{name}_valid"]) and not valid.all()
    # index-valued stages from the reference's own fp32 boxes and scores
    comp = g[f"{name}_boxes"][valid]
    order = oracle.sort_desc(g[f"{name}_score"][valid])[:pre_k]
    assert np.array_equal(order, g[f"{name}_topk_idx"])
    keep = oracle.nms(comp[order], -np.arange(len(order), dtype=np.float32), 0.7)[:post_k]
    assert np.array_equal(keep, g[f"{name}_keep"])
    assert np.array_equal(comp[order][keep], g[f"{name}_rois"])


@pytest.mark.parametrize("tag,shapes", [("hw", [(256, 320)] * 2), ("wh_like_reference", [(320, 256)] * 2)])
def test_multiscale_roi_align_vs_torchvision(oracle, tag, shapes):
    """MultiScaleRoIAlign(['0'..'3'], 7, 2) as called at models/new_model.py:127,143 (the reference passes (w, h))."""
    g = golden("msroialign")
    feats, rois5 = synth.pyramid_inputs(image_hw=(256, 320))
    scales, k_min, k_max = oracle.infer_scales([f.shape for f in feats], shapes)
    assert np.array_equal(np.asarray(scales), g[f"{tag}_scales"])
    out, lv = oracle.multiscale_roi_align(feats, rois5, shapes)
    assert np.array_equal(lv, g[f"{tag}_levels"].astype(np.int64))          # incl. the boxes exactly on a level boundary
    assert np.array_equal(out, g[f"{tag}_out"])


@pytest.mark.parametrize("name,kw", [("voc", dict(seed=8300, N=20646, S=128, C=21)),
                                     ("coco", dict(seed=8301, N=37350, S=128, C=81, n_pos=128, n_neg=128)),
                                     ("small", dict(seed=8302, N=1440, S=128, C=21, n_pos=3, n_neg=253, frc_pos=5))])
def test_region_loss_vs_reference(oracle, name, kw):
    """losses/loss.py FRCNNLoss (+ models/model.py:340-341 gather) run by tests/golden/make_golden_loss.py."""
    g = golden("loss")
    x = synth.loss_inputs(**kw)
    got = oracle.region_loss(x["rpn_cls"], x["rpn_reg"], x["rpn_tcls"], x["rpn_treg"], x["frc_cls"], x["frc_reg"],
                             x["frc_tcls"], x["frc_treg"])
    np.testing.assert_allclose(got, g[f"{name}_loss"], rtol=2e-6)


# ------------------------------------------------------------------------------------ FPN-variant anchors
FPN_ANCHOR_SIZES = ["128x192", "160x160", "800x1344", "800x1333", "600x1000", "97x131"]


def _pyr_hws(h, w):
    return [(-(-h // s), -(-w // s)) for s in (4, 8, 16, 32, 64)]


def oracle_mod():
    from oracle import region_oracle
    return region_oracle


def test_tv_anchor_base_oracle_and_abi_host_function():
    """torchvision AnchorGenerator.generate_anchors (models/new_model.py:23-25): the oracle restatement and the host-side
    C-ABI function both reproduce torchvision's five base tables bit for bit."""
    from faster_rcnn_pytorch_b200 import ops
    g = golden("anchors_fpn")
    for lvl, size in enumerate((32, 64, 128, 256, 512)):
        assert np.array_equal(oracle_mod().tv_anchor_base(size), g["base"][lvl])
        assert np.array_equal(ops.tv_anchor_base(size), g["base"][lvl])


@pytest.mark.parametrize("key", FPN_ANCHOR_SIZES)
def test_tv_anchors_pyramid_oracle_matches_reference(key):
    """models/new_model.py:43-44 (AnchorGenerator on an ImageList, / (w,h,w,h)): sha256 of the reference's own array."""
    g = golden("anchors_fpn")
    h, w = map(int, key.split("x"))
    a = oracle_mod().tv_anchors_pyramid(_pyr_hws(h, w), (h, w))
    assert a.shape[0] == int(g[f"n_{key}"])
    assert np.array_equal(sha(a), g[f"sha_{key}"])
    assert np.array_equal(a[:6], g[f"head_{key}"]) and np.array_equal(a[-6:], g[f"tail_{key}"])
