"""GPU parity of the COMPOSED region stage: one call chain logits -> proposals -> feature-map rois -> RoIPool -> per-class
decode -> per-class NMS -> packed detections (FRCNN.predict, models/model.py:346-402) and the training-side chain
(models/model.py:310-335), against the oracle chain on the same inputs -- eager and from CUDA graphs.

The chain crosses two float-tolerance points (the exp() in the RPN decode and in the per-class decode, 1e-5 relative);
every index-valued stage behind them is checked bit-exactly on the fp32 values the GPU produced, and the pure oracle
chain (CPU floats end to end) must agree with the GPU chain up to near-threshold flips."""
import numpy as np
import pytest
import torch

from faster_rcnn_pytorch_b200 import ops, region, synth, targets

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _oracle_predict_from_gpu_floats(oracle, plan, feat, head_cls, head_reg, out_pool, out_det, image):
    """The reference chain for one image, every stage fed the GPU's fp32 outputs of the previous float stage."""
    it = {k: v[image].cpu().numpy() for k, v in plan.proposal.intermediates().items()}
    boxes, valid, sc = it["boxes"], it["valid"].astype(bool), it["scores"]
    src = np.nonzero(valid)[0]
    order = src[oracle.sort_desc(sc[valid])[:plan.proposal.pre_k]]
    tb = boxes[order]
    keep = oracle.nms(tb, -np.arange(len(order), dtype=np.float32), 0.7)[:plan.R]
    rois = tb[keep]                                                            # models/model.py:53-56
    pooled, rois_g, cnt = out_pool[:3]
    c = int(cnt[image])
    assert c == len(keep)
    assert np.array_equal(rois_g[image, :c].cpu().numpy(), rois)
    r5 = oracle.scale_rois(rois, plan.fhw[0], plan.fhw[1], batch_index=0)      # models/model.py:104-110
    assert np.array_equal(plan.rois5[image * plan.R:image * plan.R + c, 1:].cpu().numpy(), r5[:, 1:])
    assert (plan.rois5[image * plan.R:image * plan.R + c, 0].cpu().numpy() == image).all()
    assert (plan.rois5[image * plan.R + c:(image + 1) * plan.R, 0].cpu().numpy() == -1).all()
    want_pool, _ = oracle.roi_pool_forward(feat[image:image + 1], r5)            # :113
    assert np.array_equal(pooled[image * plan.R:image * plan.R + c].cpu().numpy(), want_pool)
    R, NC = plan.R, plan.NC
    hc = head_cls[image * R:image * R + c]
    hr = head_reg[image * R:image * R + c]
    want_prob, want_boxes = oracle.decode_classwise(hc, hr, rois, NC)           # :369-378
    prob = out_det["prob"].reshape(-1, R, NC)[image, :c].cpu().numpy()
    bx = out_det["boxes"].reshape(-1, R, 4 * NC)[image, :c].cpu().numpy()
    np.testing.assert_allclose(prob, want_prob, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(bx, want_boxes, rtol=RTOL, atol=1e-6)
    wb, wl, ws = oracle.suppress(bx, prob, NC, 0.05)                            # :382-402 on the GPU's fp32 values
    db, dl, ds, dc = out_det["det"]
    k = int(dc[image])
    assert k == len(wl)
    assert np.array_equal(db[image, :k].cpu().numpy(), wb)
    assert np.array_equal(dl[image, :k].cpu().numpy(), wl)
    assert np.array_equal(ds[image, :k].cpu().numpy(), ws)
    packed, pc = out_det["packed"], out_det["count"]
    m = min(k, plan.max_det)
    assert int(pc[image]) == m
    row = packed[image, :m].cpu().numpy()
    assert np.array_equal(row[:, :4], wb[:m]) and np.array_equal(row[:, 4], ws[:m]) and np.array_equal(row[:, 5], wl[:m].astype(np.float32))
    return dict(rois=rois, det=(wb, wl, ws))


@pytest.mark.parametrize("hw,B,NC,seed", [((600, 1000), 1, 21, 1000), ((320, 480), 3, 81, 1100)])
def test_predict_chain_eager_and_graph(oracle, hw, B, NC, seed):
    """configs[0] (one 600x1000 VOC image, 21 classes) and a small 3-image / 81-class batch."""
    C = 24
    plan = region.InferPlan(B, hw, NC, DEV, logits=True)
    fh, fw = plan.fhw
    ins = [synth.rpn_head_outputs(seed + i, hw) for i in range(B)]
    lg = np.stack([x[0] for x in ins]); rg = np.stack([x[1] for x in ins])
    feat = synth.features(seed + 7, B, C, fh, fw)
    heads = [synth.head_outputs(seed + 20 + i, plan.R, NC) for i in range(B)]
    hcls = np.concatenate([h[0] for h in heads]); hreg = np.concatenate([h[1] for h in heads])
    d_feat, d_lg, d_rg, d_hc, d_hr = dev(feat), dev(lg), dev(rg), dev(hcls), dev(hreg)

    out_pool = plan.pool(d_feat, d_lg, d_rg)
    out_det = plan.detect(d_hc, d_hr, return_all=True)
    torch.cuda.synchronize()
    anchor = oracle.enumerate_anchors(hw)
    for i in range(B):
        got = _oracle_predict_from_gpu_floats(oracle, plan, feat, hcls, hreg, out_pool, out_det, i)
        # the pure oracle chain (CPU floats from the logits on): same proposals / detections up to near-threshold flips
        ref = oracle.region_proposal(lg[i], rg[i], anchor, "test")
        n = min(len(ref["rois"]), len(got["rois"]))
        same = np.isclose(ref["rois"][:n], got["rois"][:n], rtol=1e-4, atol=1e-6).all(axis=1).mean()
        assert same >= 0.97 and abs(len(ref["rois"]) - len(got["rois"])) <= 3

    # the same two halves from CUDA graphs, on NEW inputs written into the captured buffers
    g_pool, o_pool = plan.capture_pool(d_feat, d_lg, d_rg)
    g_det, o_det = _capture_detect_all(plan, d_hc, d_hr)
    ins2 = [synth.rpn_head_outputs(seed + 500 + i, hw) for i in range(B)]
    lg2 = np.stack([x[0] for x in ins2]); rg2 = np.stack([x[1] for x in ins2])
    feat2 = synth.features(seed + 507, B, C, fh, fw)
    d_feat.copy_(dev(feat2)); d_lg.copy_(dev(lg2)); d_rg.copy_(dev(rg2))
    g_pool.replay(); g_det.replay()
    torch.cuda.synchronize()
    for i in range(B):
        _oracle_predict_from_gpu_floats(oracle, plan, feat2, hcls, hreg, o_pool, o_det, i)


def _capture_detect_all(plan, d_hc, d_hr):
    return region._capture(lambda: plan.detect(d_hc, d_hr, return_all=True), plan.device)


def test_train_chain_eager_and_graph(oracle):
    """configs[2]-shaped chain for 3 images: both target makers (device replay of torch's randperm stream), RoIPool of the
    sampled rois, RoIPool backward -- against the oracle chain driven by the same mt19937 stream."""
    hw, B, G, R, C = (600, 1000), 3, 8, 2000, 16
    fh, fw = hw[0] // 16, hw[1] // 16
    anchor = oracle.enumerate_anchors(hw)
    gts = np.stack([synth.gt_boxes(6000 + i, G)[0] for i in range(B)])
    labs = np.stack([synth.gt_boxes(6000 + i, G)[1] for i in range(B)])
    props = np.stack([synth.random_boxes(6100 + i, R)[0] for i in range(B)])
    feat = synth.features(6300, B, C, fh, fw)
    gout = np.random.RandomState(6400).standard_normal((B * 128, C, 7, 7)).astype(np.float32)
    d_feat, d_gt, d_lab, d_props, d_gout = dev(feat), dev(gts), dev(labs), dev(props), dev(gout)
    pcnt = torch.full((B,), R, dtype=torch.int32, device=DEV)

    def check(t, pooled, gin, seed):
        rp = oracle.HostRandperm(seed)
        want_gin = np.zeros_like(feat)
        for i in range(B):
            wr = oracle.rpn_targets(gts[i], anchor, rp)
            wf = oracle.frcnn_targets(gts[i], labs[i], props[i], rp)
            assert np.array_equal(t["rpn_cls"][i].cpu().numpy(), wr["labels"])
            np.testing.assert_allclose(t["rpn_reg"][i].cpu().numpy(), wr["reg"], rtol=RTOL, atol=1e-5)
            assert np.array_equal(t["frcnn_cls"][i].cpu().numpy(), wf["cls"])
            assert np.array_equal(t["sample_rois"][i].cpu().numpy(), wf["sample_rois"])
            r5 = oracle.scale_rois(wf["sample_rois"], fh, fw, batch_index=0)
            wo, wa = oracle.roi_pool_forward(feat[i:i + 1], r5)
            assert np.array_equal(pooled[i * 128:(i + 1) * 128].cpu().numpy(), wo)
            want_gin[i] = oracle.roi_pool_backward(gout[i * 128:(i + 1) * 128], wa, r5, (1, C, fh, fw))[0]
        g = gin.cpu().numpy()
        assert np.abs(g - want_gin).max() <= 1e-5 * np.abs(want_gin).max()

    torch.manual_seed(6200)
    plan = region.TrainPlan(B, hw, DEV)
    t, pooled = plan.targets_and_pool(d_feat, d_gt, d_lab, d_props, pcnt)
    gin = plan.pool_backward(d_gout)
    torch.cuda.synchronize()
    check(t, pooled, gin, 6200)

    torch.manual_seed(6201)
    plan2 = region.TrainPlan(B, hw, DEV)
    state0 = plan2.generator.state.clone()
    g_fwd, (t2, pooled2) = plan2.capture_targets_and_pool(d_feat, d_gt, d_lab, d_props, pcnt)
    g_bwd, gin2 = plan2.capture_pool_backward(d_gout)
    plan2.generator.state.copy_(state0)            # the warm-up + capture runs advanced the stream: rewind, then replay
    g_fwd.replay(); g_bwd.replay()
    torch.cuda.synchronize()
    check(t2, pooled2, gin2, 6201)
