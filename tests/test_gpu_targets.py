"""GPU parity: RPN / Fast R-CNN target makers (C ABI assign + finalize kernels around the host randperm)
vs the reference goldens (torch.manual_seed + torch.randperm) and the CPU oracle.
Labels, argmax matches and sampled indices bit-exact; encode() targets within 1e-5 (logf)."""
import numpy as np
import pytest
import torch

from conftest import golden
from faster_rcnn_pytorch_b200 import ops, synth, targets

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5

CASES = [("kat6", (600, 1000), None, 2, 7), ("c3_0", (600, 1000), 3000, 8, 3000), ("c3_1", (600, 1000), 3001, 8, 3001),
         ("many_gt", (600, 1000), 3100, 160, 3100), ("one_gt", (320, 480), 3200, 1, 3200), ("small", (160, 256), 3300, 3, 3300)]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def close(a, b, atol=1e-6):
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=atol)


def _gt(gseed, G):
    if gseed is None:
        return np.array([[0.1, 0.2, 0.5, 0.7], [0.3, 0.3, 0.9, 0.95]], np.float32), np.array([11, 14], np.int64)
    return synth.gt_boxes(gseed, G)


@pytest.mark.parametrize("name,hw,gseed,G,tseed", CASES)
@pytest.mark.parametrize("gen_anchors", [True, False])
def test_rpn_targets_reference_goldens(oracle, name, hw, gseed, G, tseed, gen_anchors):
    g = golden("targets")
    gt, _ = _gt(gseed, G)
    torch.manual_seed(tseed)        # the reference's sampling stream (models/model.py:228,235)
    if gen_anchors:
        labels, reg = targets.rpn_targets(dev(gt[None]), None, image_hw=hw)
    else:
        labels, reg = targets.rpn_targets(dev(gt[None]), None, anchors=dev(oracle.enumerate_anchors(hw)))
    labels, reg = labels[0].cpu().numpy(), reg[0].cpu().numpy()
    assert labels.dtype == np.int64
    assert np.array_equal(labels, g[f"{name}_rpn_cls"].astype(np.int64))       # bit-exact labels incl. sampling
    if f"{name}_rpn_reg" in g:
        close(reg, g[f"{name}_rpn_reg"])
    else:
        close(reg[labels >= 0], g[f"{name}_rpn_reg_sampled"])
    if name == "kat6":
        assert ((labels == 1).sum(), (labels == 0).sum(), (labels == -1).sum()) == (41, 215, 20390)


@pytest.mark.parametrize("hw,G", [((600, 1000), 8), ((800, 1333), 8), ((608, 1008), 40)])
def test_rpn_targets_batched_ragged_gt_vs_oracle(oracle, hw, G):
    """Batch of images with different GT counts (padded to Gmax) -> per-image parity with the oracle,
    including the intermediate IoU row max / argmax (bit-exact fp32)."""
    B = 5
    anchor = oracle.enumerate_anchors(hw)
    N = anchor.shape[0]
    gts, cnt = [], []
    for i in range(B):
        gi = max(1, G - 3 * i)
        gts.append(synth.gt_boxes(4000 + i, gi)[0]); cnt.append(gi)
    pad = np.zeros((B, G, 4), np.float32)
    for i, x in enumerate(gts):
        pad[i, :cnt[i]] = x
    rp = oracle.HostRandperm(4100)
    ws = ops.rpn_targets_assign(dev(pad), dev(np.asarray(cnt, np.int32)), N, image_hw=hw)
    want = [oracle.rpn_targets(gts[i], anchor, oracle.HostRandperm(4100 + i)) for i in range(B)]
    im, am, counts = ws["iou_max"].cpu().numpy(), ws["argmax"].cpu().numpy(), ws["counts"].cpu().numpy()
    for i in range(B):
        ins = want[i]["inside"]
        assert np.array_equal(im[i][ins], want[i]["iou_max"][ins])               # fp32 IoU bit-exact
        assert np.array_equal(am[i][ins], want[i]["argmax"][ins])
        assert tuple(counts[i]) == (want[i]["n_pos"], want[i]["n_neg"])
    # full path, each image with its own host stream
    for i in range(B):
        labels, reg = targets.rpn_targets(dev(pad[i:i + 1]), dev(np.asarray(cnt[i:i + 1], np.int32)), image_hw=hw,
                                          randperm=oracle.HostRandperm(4100 + i))
        assert np.array_equal(labels[0].cpu().numpy(), want[i]["labels"])
        close(reg[0].cpu().numpy(), want[i]["reg"])
    # whole batch through one call with one shared stream == images processed in order on that stream
    labels, reg = targets.rpn_targets(dev(pad), dev(np.asarray(cnt, np.int32)), image_hw=hw, randperm=rp)
    rp2 = oracle.HostRandperm(4100)
    for i in range(B):
        w = oracle.rpn_targets(gts[i], anchor, rp2)
        assert np.array_equal(labels[i].cpu().numpy(), w["labels"])


def test_rpn_targets_edge_cases(oracle):
    hw = (160, 256)
    anchor = oracle.enumerate_anchors(hw)
    # a GT that overlaps no anchor well still yields >= 1 positive (low-quality match, :206-213)
    tiny = np.array([[0.5, 0.5, 0.5005, 0.5005]], np.float32)
    labels, _ = targets.rpn_targets(dev(tiny[None]), None, image_hw=hw, randperm=oracle.HostRandperm(1))
    w = oracle.rpn_targets(tiny, anchor, oracle.HostRandperm(1))
    assert np.array_equal(labels[0].cpu().numpy(), w["labels"]) and (w["labels"] == 1).sum() >= 1
    # duplicate GT boxes: per-GT argmax ties resolve to the first anchor index for both
    dup = np.array([[0.2, 0.2, 0.6, 0.7], [0.2, 0.2, 0.6, 0.7]], np.float32)
    labels, reg = targets.rpn_targets(dev(dup[None]), None, image_hw=hw, randperm=oracle.HostRandperm(2))
    w = oracle.rpn_targets(dup, anchor, oracle.HostRandperm(2))
    assert np.array_equal(labels[0].cpu().numpy(), w["labels"])
    close(reg[0].cpu().numpy(), w["reg"])
    # G = 0 raises like the reference (models/model.py:199)
    with pytest.raises(IndexError):
        targets.rpn_targets(torch.zeros((1, 0, 4), device=DEV), None, image_hw=hw)


@pytest.mark.parametrize("name,hw,gseed,G,tseed", CASES)
def test_frcnn_targets_reference_goldens(oracle, name, hw, gseed, G, tseed):
    g = golden("targets")
    gt, lab = _gt(gseed, G)
    rois, _ = synth.random_boxes(tseed + 50, 2000)
    torch.manual_seed(tseed + 1)
    cls, reg, srois, kidx, n = targets.frcnn_targets(dev(rois[None]), None, dev(gt[None]), None, dev(lab[None]))
    assert int(n[0]) == 128
    assert np.array_equal(cls[0].cpu().numpy(), g[f"{name}_frcnn_cls"].astype(np.int64))   # sampled indices bit-exact
    assert np.array_equal(srois[0].cpu().numpy(), g[f"{name}_frcnn_rois"])
    close(reg[0].cpu().numpy(), g[f"{name}_frcnn_reg"], atol=2e-5)
    w = oracle.frcnn_targets(gt, lab, rois, oracle.HostRandperm(tseed + 1))
    assert np.array_equal(kidx[0].cpu().numpy().astype(np.int64), w["keep"])


def test_frcnn_targets_batched_ragged_and_short(oracle):
    B, R, G = 4, 2000, 8
    rois = np.zeros((B, R, 4), np.float32); gts = np.zeros((B, G, 4), np.float32); labs = np.zeros((B, G), np.int64)
    rc = np.asarray([2000, 1500, 40, 3], np.int32); gc = np.asarray([8, 5, 2, 1], np.int32)
    for i in range(B):
        rois[i, :rc[i]] = synth.random_boxes(5000 + i, int(rc[i]))[0]
        b, l = synth.gt_boxes(5100 + i, int(gc[i]))
        gts[i, :gc[i]] = b; labs[i, :gc[i]] = l
    rp = oracle.HostRandperm(5200)
    cls, reg, srois, kidx, n = targets.frcnn_targets(dev(rois), dev(rc), dev(gts), dev(gc), dev(labs), randperm=rp)
    rp2 = oracle.HostRandperm(5200)
    for i in range(B):
        w = oracle.frcnn_targets(gts[i, :gc[i]], labs[i, :gc[i]], rois[i, :rc[i]], rp2)
        k = len(w["keep"])
        assert int(n[i]) == k       # fewer than 128 when there are too few negatives (SURVEY §8a edge case)
        assert np.array_equal(cls[i, :k].cpu().numpy(), w["cls"])
        assert np.array_equal(kidx[i, :k].cpu().numpy().astype(np.int64), w["keep"])
        assert np.array_equal(srois[i, :k].cpu().numpy(), w["sample_rois"])
        close(reg[i, :k].cpu().numpy(), w["reg"], atol=2e-5)
        assert (cls[i, k:].cpu().numpy() == -1).all()
    assert int(n[3]) < 128


def test_make_targets_single_sync_matches_sequential(oracle):
    """targets.make_targets: both makers, one D2H for the batch, draws in the reference's per-image order."""
    hw, B, G, R = (600, 1000), 3, 8, 2000
    anchor = oracle.enumerate_anchors(hw)
    gts = np.stack([synth.gt_boxes(6000 + i, G)[0] for i in range(B)])
    labs = np.stack([synth.gt_boxes(6000 + i, G)[1] for i in range(B)])
    rois = np.stack([synth.random_boxes(6100 + i, R)[0] for i in range(B)])
    out = targets.make_targets(dev(gts), None, dev(labs), dev(rois), None, image_hw=hw, randperm=oracle.HostRandperm(6200))
    rp = oracle.HostRandperm(6200)
    for i in range(B):
        wr = oracle.rpn_targets(gts[i], anchor, rp)
        wf = oracle.frcnn_targets(gts[i], labs[i], rois[i], rp)
        assert np.array_equal(out["rpn_cls"][i].cpu().numpy(), wr["labels"])
        close(out["rpn_reg"][i].cpu().numpy(), wr["reg"])
        assert np.array_equal(out["frcnn_cls"][i].cpu().numpy(), wf["cls"])
        assert np.array_equal(out["sample_rois"][i].cpu().numpy(), wf["sample_rois"])
        close(out["frcnn_reg"][i].cpu().numpy(), wf["reg"], atol=2e-5)


# ------------------------------------------------------------------------------------ FPN-variant target makers
FPN_CASES = [("a", (320, 480), 7000, 5, 7000), ("b", (600, 1000), 7001, 8, 7001), ("many", (320, 480), 7002, 150, 7002),
             ("one", (160, 256), 7003, 1, 7003)]


@pytest.mark.parametrize("name,hw,gseed,G,tseed", FPN_CASES)
def test_fpn_variant_targets_reference_goldens(oracle, name, hw, gseed, G, tseed):
    """models/new_model.py:153-206,299-349 through the same kernels with the variant knobs (eps-free IoU, no inside
    filter, tie-inclusive low-quality match, 512 samples / <= 128 positives, labels used as they are)."""
    g = golden("targets_fpn")
    anchors = oracle.enumerate_anchors(hw)
    gt, lab = synth.gt_boxes(gseed, G)
    torch.manual_seed(tseed)
    labels, reg = targets.rpn_targets(dev(gt[None]), None, anchors=dev(anchors), variant="fpn")
    labels, reg = labels[0].cpu().numpy(), reg[0].cpu().numpy()
    assert np.array_equal(labels, g[f"{name}_rpn_cls"].astype(np.int64))
    close(reg[labels >= 0], g[f"{name}_rpn_reg_sampled"])
    rois, _ = synth.random_boxes(tseed + 50, 2000)
    torch.manual_seed(tseed + 1)
    cls, freg, srois, kidx, n = targets.frcnn_targets(dev(rois[None]), None, dev(gt[None]), None, dev((lab + 1)[None]),
                                                      variant="fpn")
    assert int(n[0]) == 512 and cls.shape == (1, 512)
    assert np.array_equal(cls[0].cpu().numpy(), g[f"{name}_frcnn_cls"].astype(np.int64))
    assert np.array_equal(srois[0].cpu().numpy(), g[f"{name}_frcnn_rois"])
    close(freg[0].cpu().numpy(), g[f"{name}_frcnn_reg"], atol=2e-5)


def test_fpn_variant_tie_inclusive_zero_iou(oracle):
    """A GT overlapping no anchor: every anchor with IoU == 0 for it becomes positive (torch.where(iou == max))."""
    g = golden("targets_fpn")
    gt = np.array([[0.2, 0.2, 0.6, 0.7], [5.0, 5.0, 5.1, 5.1]], np.float32)
    torch.manual_seed(9)
    labels, _ = targets.rpn_targets(dev(gt[None]), None, anchors=dev(g["far_anchors"]), variant="fpn")
    assert np.array_equal(labels[0].cpu().numpy(), g["far_rpn_cls"].astype(np.int64))


# ------------------------------------------------------------------------------------ device-side sampling
@pytest.mark.parametrize("name,hw,gseed,G,tseed", CASES)
def test_device_sampling_reference_goldens(name, hw, gseed, G, tseed):
    """frr_sample_targets: torch's mt19937 state copied to the device once, the randperm draws of
    models/model.py:149,155,228,235 replayed there -- same goldens as the host path, no synchronisation in between."""
    g = golden("targets")
    gt, lab = _gt(gseed, G)
    torch.manual_seed(tseed)
    gen = targets.DeviceGenerator(DEV)
    labels, reg = targets.rpn_targets(dev(gt[None]), None, image_hw=hw, generator=gen)
    labels = labels[0].cpu().numpy()
    assert np.array_equal(labels, g[f"{name}_rpn_cls"].astype(np.int64))
    rois, _ = synth.random_boxes(tseed + 50, 2000)
    torch.manual_seed(tseed + 1)
    gen.sync_from_torch()
    cls, freg, srois, kidx, n = targets.frcnn_targets(dev(rois[None]), None, dev(gt[None]), None, dev(lab[None]), generator=gen)
    assert n.is_cuda and int(n[0]) == 128
    assert np.array_equal(cls[0].cpu().numpy(), g[f"{name}_frcnn_cls"].astype(np.int64))
    assert np.array_equal(srois[0].cpu().numpy(), g[f"{name}_frcnn_rois"])
    close(freg[0].cpu().numpy(), g[f"{name}_frcnn_reg"], atol=2e-5)


@pytest.mark.parametrize("name,hw,gseed,G,tseed", FPN_CASES)
def test_device_sampling_fpn_variant_goldens(oracle, name, hw, gseed, G, tseed):
    g = golden("targets_fpn")
    anchors = oracle.enumerate_anchors(hw)
    gt, lab = synth.gt_boxes(gseed, G)
    torch.manual_seed(tseed)
    gen = targets.DeviceGenerator(DEV)
    labels, _ = targets.rpn_targets(dev(gt[None]), None, anchors=dev(anchors), variant="fpn", generator=gen)
    assert np.array_equal(labels[0].cpu().numpy(), g[f"{name}_rpn_cls"].astype(np.int64))
    rois, _ = synth.random_boxes(tseed + 50, 2000)
    torch.manual_seed(tseed + 1)
    gen.sync_from_torch()
    cls, _, srois, _, n = targets.frcnn_targets(dev(rois[None]), None, dev(gt[None]), None, dev((lab + 1)[None]),
                                                variant="fpn", generator=gen)
    assert int(n[0]) == 512
    assert np.array_equal(cls[0].cpu().numpy(), g[f"{name}_frcnn_cls"].astype(np.int64))
    assert np.array_equal(srois[0].cpu().numpy(), g[f"{name}_frcnn_rois"])


def test_device_sampling_batch_equals_host_stream_and_state():
    """Two consecutive batches (crowded image with > 128 positives, ragged GT, an image with too few negatives) through
    make_targets: the device generator yields bit-identical targets AND leaves torch's generator, after
    sync_to_torch(), exactly where the host path leaves it."""
    hw, B, G, R = (600, 1000), 6, 160, 2000
    gts = np.zeros((B, G, 4), np.float32); labs = np.zeros((B, G), np.int64)
    gc = np.asarray([8, 160, 1, 40, 3, 8], np.int32)
    rc = np.asarray([2000, 2000, 1500, 2000, 20, 700], np.int32)
    rois = np.zeros((B, R, 4), np.float32)
    for i in range(B):
        b, l = synth.gt_boxes(9000 + i, int(gc[i]))
        gts[i, :gc[i]] = b; labs[i, :gc[i]] = l
        rois[i, :rc[i]] = synth.random_boxes(9100 + i, int(rc[i]))[0]
    args = (dev(gts), dev(gc), dev(labs), dev(rois), dev(rc))

    torch.manual_seed(777)
    host = [targets.make_targets(*args, image_hw=hw) for _ in range(2)]
    host_state = torch.get_rng_state().clone()
    host_next = torch.randperm(1000)

    torch.manual_seed(777)
    gen = targets.DeviceGenerator(DEV)
    devo = [targets.make_targets(*args, image_hw=hw, generator=gen) for _ in range(2)]   # no sync between the batches
    gen.sync_to_torch()
    assert torch.equal(torch.get_rng_state(), host_state)
    assert torch.equal(torch.randperm(1000), host_next)
    for h, d in zip(host, devo):
        for k in ("rpn_cls", "rpn_reg", "frcnn_cls", "frcnn_reg", "sample_rois", "keep_index"):
            assert torch.equal(h[k], d[k]), k
        assert np.array_equal(d["n_samples"].cpu().numpy(), h["n_samples"])
    assert (host[0]["rpn_cls"][1] == 1).sum() == 128          # the crowded image exercised the positive draw
    assert int(host[0]["n_samples"][4]) < 128                 # too few negatives -> short sample


def test_device_generator_round_trip_without_draws():
    torch.manual_seed(5)
    before = torch.get_rng_state().clone()
    want = torch.randperm(50)
    torch.set_rng_state(before)
    targets.DeviceGenerator(DEV).sync_to_torch()              # freshly seeded: left = 1 <-> position 624
    assert torch.equal(torch.randperm(50), want)


def test_device_sampling_more_than_256_images():
    """frr_sample_targets walks its job table in shared memory 256 images at a time: a batch of 300 goes through two
    stream-ordered slices and must still equal the host stream image by image."""
    hw, B, G, R = (160, 256), 300, 3, 200
    gts = np.stack([synth.gt_boxes(9500 + i, G)[0] for i in range(B)])
    labs = np.stack([synth.gt_boxes(9500 + i, G)[1] for i in range(B)])
    rois = np.stack([synth.random_boxes(9900 + (i % 7), R)[0] for i in range(B)])
    args = (dev(gts), None, dev(labs), dev(rois), None)
    torch.manual_seed(4242)
    host = targets.make_targets(*args, image_hw=hw)
    host_state = torch.get_rng_state().clone()
    torch.manual_seed(4242)
    gen = targets.DeviceGenerator(DEV)
    devo = targets.make_targets(*args, image_hw=hw, generator=gen)
    gen.sync_to_torch()
    assert torch.equal(torch.get_rng_state(), host_state)
    for k in ("rpn_cls", "frcnn_cls", "keep_index", "sample_rois"):
        assert torch.equal(host[k], devo[k]), k
