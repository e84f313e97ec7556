"""Developer tool (GPU box): per-kernel timings of the training / inference region-stage kernels at the
BASELINE config-3 / config-4 shapes, with algorithmic bytes and the fraction of the measured HBM
peak; torchvision's own sm_100 CUDA kernels are timed beside them on the same inputs.

    python tools/stage_profile.py [--quick]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from faster_rcnn_pytorch_b200 import ops, synth, targets, region

dev = torch.device("cuda:0")
PEAK = 6538.3
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=30, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()            # > L2: evict
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3  # us


def report(name, us, nbytes=None, **kw):
    d = {"kernel": name, "us": round(us, 2)}
    if nbytes:
        d["MB"] = round(nbytes / 1e6, 2)
        d["GBps"] = round(nbytes / us / 1e3, 1)
        d["hbm_frac"] = round(nbytes / us / 1e3 / PEAK, 3)
    d.update(kw)
    print(json.dumps(d), flush=True)


def nms_race(flush):
    """torchvision's sm_100 CUDA `nms` (IoU bitmask [n, n/64] + gather_keep_from_mask) against
    frr_nms_sorted on the same score-sorted RPN boxes (models/model.py:53: 12000 -> 2000 train, 6000 -> 300 test) and
    torchvision `batched_nms` against frr_class_nms for the 300 x 80 per-class case (models/model.py:394); also
    where the keep lists differ (torchvision's CUDA kernel narrows the threshold to fp32 and contracts Sa + Sb)."""
    import torchvision
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import nms_cases
    for tag, n, post, B in [("train_12000_2000", 12000, 2000, 64), ("test_6000_300", 6000, 300, 8)]:
        boxes = np.stack([nms_cases.rpn_like(2000 + i, n) for i in range(B)])
        d = torch.from_numpy(boxes).to(dev)
        sc = torch.arange(n, 0, -1, device=dev, dtype=torch.float32)
        tv1 = timeit(lambda: torchvision.ops.nms(d[0], sc, 0.7)[:post], flush=flush)
        report(f"nms/{tag}/torchvision_cuda_nms_one_image", tv1, n=n)

        def tv_loop():
            return [torchvision.ops.nms(d[i], sc, 0.7)[:post] for i in range(B)]
        tvB = timeit(tv_loop, reps=5, warm=2, flush=flush)
        report(f"nms/{tag}/torchvision_cuda_nms_loop_over_{B}_images", tvB, per_image_us=round(tvB / B, 2))
        idxs = torch.arange(B, device=dev).repeat_interleave(n)

        def tv_batched():
            return torchvision.ops.batched_nms(d.reshape(-1, 4), sc.repeat(B), idxs, 0.7)
        tvb = timeit(tv_batched, reps=5, warm=2, flush=flush)
        report(f"nms/{tag}/torchvision_cuda_batched_nms_{B}_images", tvb, per_image_us=round(tvb / B, 2))
        one = d[:1].contiguous()
        us1 = timeit(lambda: ops.nms_sorted(one, 0.7, max_keep=post, unit_boxes=True), flush=flush)
        report(f"nms/{tag}/frr_nms_sorted_one_image", us1, variant=ops.nms_variant(1, n, 0.7, post, unit_boxes=True))
        usB = timeit(lambda: ops.nms_sorted(d, 0.7, max_keep=post, unit_boxes=True), flush=flush)
        report(f"nms/{tag}/frr_nms_sorted_{B}_images", usB, per_image_us=round(usB / B, 2),
               variant=ops.nms_variant(B, n, 0.7, post, unit_boxes=True),
               speedup_vs_torchvision_loop=round(tvB / usB, 1))
        keep, cnt, _ = ops.nms_sorted(d, 0.7, max_keep=post, unit_boxes=True)
        diff = 0
        for i in range(B):
            a = torchvision.ops.nms(d[i], sc, 0.7)[:post].cpu().numpy()
            b = keep[i, :int(cnt[i])].cpu().numpy()
            diff += int(len(a) != len(b) or not np.array_equal(a, b))
        report(f"nms/{tag}/images_where_torchvision_cuda_keep_list_differs_from_cpu_exact", 0.0, images=B, differing=diff)
    # per-class NMS: 8 images x 300 rois x 80 classes, thres 0.05 / IoU 0.3
    B, R, C = 8, 300, 81
    cls = torch.from_numpy(np.stack([synth.head_outputs(600 + i, R, C)[0] for i in range(B)])).to(dev)
    reg = torch.from_numpy(np.stack([synth.head_outputs(600 + i, R, C)[1] for i in range(B)])).to(dev)
    rr = torch.from_numpy(np.stack([synth.random_boxes(700 + i, R)[0] for i in range(B)])).to(dev)
    prob, boxes = ops.decode_classwise(cls.reshape(B * R, C), reg.reshape(B * R, 4 * C), rr.reshape(B * R, 4), C)
    prob = prob.reshape(B, R, C)
    boxes = boxes.reshape(B, R, C, 4)

    def tv_class():
        # the reference's loop (models/model.py:386-398) collapsed into ONE batched_nms over (image, class) groups
        m = prob[:, :, 1:] > 0.05
        bi, ri, ci = torch.nonzero(m, as_tuple=True)
        return torchvision.ops.batched_nms(boxes[bi, ri, ci + 1], prob[bi, ri, ci + 1], bi * C + ci, 0.3)
    us_tv = timeit(tv_class, reps=10, flush=flush)
    us = timeit(lambda: ops.class_nms(prob, boxes.reshape(B, R, 4 * C), C, score_thres=0.05), flush=flush)
    report("nms/class_300x80x8/torchvision_cuda_batched_nms(mask+nonzero+gather+nms)", us_tv)
    report("nms/class_300x80x8/frr_class_nms", us, speedup=round(us_tv / us, 1))


def main():
    quick = "--quick" in sys.argv
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    import torchvision
    if "--skip-nms" not in sys.argv:
        nms_race(flush)
    if "--only-nms" in sys.argv:
        return
    for tag, B, C, fh, fw, per_img, real in [("cfg3_train", 16, 512, 37, 62, 128, False), ("cfg3_train_rpnrois", 16, 512, 37, 62, 128, True),
                                              ("cfg4_infer", 8, 512, 50, 83, 300, False), ("cfg4_infer_rpnrois", 8, 512, 50, 83, 300, True)]:
        K = B * per_img
        feat = torch.from_numpy(synth.features(1, B, C, fh, fw)).to(dev)
        if real:   # proposals of the RPN pipeline on synthetic head outputs: first per_img per image
            hw = (fh * 16, fw * 16)
            n = synth.num_anchors(hw)
            rs = np.random.RandomState(77)
            lg = torch.from_numpy(rs.standard_normal((B, n, 2)).astype(np.float32)).to(dev)
            rg = torch.from_numpy((rs.standard_normal((B, n, 4)) * 0.2).astype(np.float32)).to(dev)
            pr, pc = region.rpn_proposals(lg, rg, image_hw=hw, mode="train")
            assert int(pc.min()) >= per_img
            boxes = pr[:, :per_img].cpu().numpy() * np.array([fw, fh, fw, fh], np.float32)
            rois_np = np.concatenate([np.concatenate([np.full((per_img, 1), b, np.float32), boxes[b]], 1) for b in range(B)])
        else:
            rois_np = np.concatenate([np.concatenate([np.full((per_img, 1), b, np.float32),
                                                      synth.random_boxes(10 + b, per_img)[0] * np.array([fw, fh, fw, fh], np.float32)], 1)
                                      for b in range(B)])
        rois = torch.from_numpy(rois_np.astype(np.float32)).to(dev)
        go = torch.randn((K, C, 7, 7), device=dev)
        fbytes = feat.numel() * 4
        obytes = K * C * 49 * 4
        for cl in (False, True):
            f = feat.contiguous(memory_format=torch.channels_last) if cl else feat
            lay = "nhwc" if cl else "nchw"
            out, arg = ops.roi_pool_forward(f, rois)
            us = timeit(lambda: ops.roi_pool_forward(f, rois), flush=flush)
            report(f"{tag}/roi_pool_fwd/{lay}", us, fbytes + 2 * obytes, K=K)
            us = timeit(lambda: ops.roi_pool_forward(f, rois, want_argmax=False), flush=flush)
            report(f"{tag}/roi_pool_fwd_noargmax/{lay}", us, fbytes + obytes, K=K)
            us = timeit(lambda: ops.roi_pool_backward(go, arg, rois, feat.shape, channels_last=cl), flush=flush)
            report(f"{tag}/roi_pool_bwd/{lay}", us, 2 * obytes + fbytes, K=K)
            us = timeit(lambda: ops.roi_align_forward(f, rois, sampling_ratio=2), flush=flush)
            report(f"{tag}/roi_align_fwd/{lay}", us, fbytes + obytes, K=K)
            us = timeit(lambda: ops.roi_align_backward(go, rois, feat.shape, sampling_ratio=2, channels_last=cl), flush=flush)
            report(f"{tag}/roi_align_bwd/{lay}", us, obytes + fbytes, K=K)
        # torchvision's CUDA kernels (generic sm_100 recompiles) on the same inputs
        us = timeit(lambda: torch.ops.torchvision.roi_pool(feat, rois, 1.0, 7, 7), flush=flush)
        report(f"{tag}/torchvision_roi_pool_fwd", us, fbytes + 2 * obytes)
        o_tv, a_tv = torch.ops.torchvision.roi_pool(feat, rois, 1.0, 7, 7)
        us = timeit(lambda: torch.ops.torchvision._roi_pool_backward(go, rois, a_tv, 1.0, 7, 7, B, C, fh, fw), flush=flush)
        report(f"{tag}/torchvision_roi_pool_bwd", us, 2 * obytes + fbytes)
        us = timeit(lambda: torch.ops.torchvision.roi_align(feat, rois, 1.0, 7, 7, 2, False), flush=flush)
        report(f"{tag}/torchvision_roi_align_fwd", us, fbytes + obytes)
        us = timeit(lambda: torch.ops.torchvision._roi_align_backward(go, rois, 1.0, 7, 7, B, C, fh, fw, 2, False), flush=flush)
        report(f"{tag}/torchvision_roi_align_bwd", us, obytes + fbytes)
        if quick and real:
            break

    # ---- target makers (config 3: B=16, G=8, 600x1000, 2000 proposals)
    hw, B, G, R = (600, 1000), 16, 8, 2000
    N = synth.num_anchors(hw)
    gt = torch.from_numpy(np.stack([synth.gt_boxes(3000 + i, G)[0] for i in range(B)])).to(dev)
    lab = torch.from_numpy(np.stack([synth.gt_boxes(3000 + i, G)[1] for i in range(B)])).to(dev)
    props = torch.from_numpy(np.stack([synth.random_boxes(3100 + i, R)[0] for i in range(B)])).to(dev)
    us = timeit(lambda: ops.rpn_targets_assign(gt, None, N, image_hw=hw), flush=flush)
    report("cfg3/rpn_targets_assign", us, B * N * (4 + 4 + 1 + 8), B=B, N=N)
    ws = ops.rpn_targets_assign(gt, None, N, image_hw=hw)
    us = timeit(lambda: ops.rpn_targets_finalize(ws, None, None), flush=flush)
    report("cfg3/rpn_targets_finalize", us, B * N * (4 + 1 + 8 + 16), B=B, N=N)
    us = timeit(lambda: ops.frcnn_targets_assign(props, None, gt, None), flush=flush)
    report("cfg3/frcnn_targets_assign", us, B * (R + G) * (16 + 17), B=B)
    torch.manual_seed(0)
    us = timeit(lambda: targets.make_targets(gt, None, lab, props, None, image_hw=hw), reps=10)
    report("cfg3/make_targets_host_roundtrip(4 kernels + 1 D2H + randperm + H2D)", us, B=B)
    gen = targets.DeviceGenerator(gt.device)
    us = timeit(lambda: targets.make_targets(gt, None, lab, props, None, image_hw=hw, generator=gen), reps=10)
    report("cfg3/make_targets_device_sampling(6 kernels, no sync)", us, B=B)
    wr = ops.rpn_targets_assign(gt, None, synth.num_anchors(hw), image_hw=hw)
    wf = ops.frcnn_targets_assign(props, None, gt, None)
    us = timeit(lambda: ops.sample_targets(gen.state, ws_rpn=wr, ws_frcnn=wf), reps=10)
    report("cfg3/sample_targets(mt19937 stream + Fisher-Yates apply)", us, B=B)

    # ---- detection post-processing (config 4: B=8, R=300, C=81)
    B, R, C = 8, 300, 81
    cls = torch.from_numpy(np.stack([synth.head_outputs(600 + i, R, C)[0] for i in range(B)])).to(dev)
    reg = torch.from_numpy(np.stack([synth.head_outputs(600 + i, R, C)[1] for i in range(B)])).to(dev)
    rr = torch.from_numpy(np.stack([synth.random_boxes(700 + i, R)[0] for i in range(B)])).to(dev)
    us = timeit(lambda: ops.decode_classwise(cls.reshape(B * R, C), reg.reshape(B * R, 4 * C), rr.reshape(B * R, 4), C), flush=flush)
    report("cfg4/decode_classwise", us, B * R * C * (4 + 16 + 4 + 16), B=B)
    prob, boxes = ops.decode_classwise(cls.reshape(B * R, C), reg.reshape(B * R, 4 * C), rr.reshape(B * R, 4), C)
    for thres in (0.05, 0.005):
        us = timeit(lambda: ops.class_nms(prob.reshape(B, R, C), boxes.reshape(B, R, 4 * C), C, score_thres=thres), flush=flush)
        dc = ops.class_nms(prob.reshape(B, R, C), boxes.reshape(B, R, 4 * C), C, score_thres=thres)[3]
        report(f"cfg4/class_nms(thres={thres})", us, B * R * C * 20, B=B, dets=int(dc.sum()))

    # ---- inference proposals (config 4: 800x1333, 6000 -> 300) and config 1 (600x1000, B=1)
    for tag, hw, B in [("cfg4", (800, 1333), 8), ("cfg1", (600, 1000), 1)]:
        n = synth.num_anchors(hw)
        rs = np.random.RandomState(1)
        lg = torch.from_numpy(rs.standard_normal((B, n, 2)).astype(np.float32)).to(dev)
        rg = torch.from_numpy((rs.standard_normal((B, n, 4)) * 0.2).astype(np.float32)).to(dev)
        plan = region.ProposalPlan(B, n, dev, image_hw=hw, mode="test")
        us = timeit(lambda: plan.run(lg, rg), flush=flush)
        report(f"{tag}/rpn_proposals_test_mode", us, B=B, N=n)


if __name__ == "__main__":
    main()
