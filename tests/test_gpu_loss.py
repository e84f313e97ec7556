"""GPU parity: the fused loss kernel (C ABI frr_region_loss) vs the reference's losses/loss.py goldens (losses and
autograd gradients, tests/golden/make_golden_loss.py) and the CPU oracle.  Floating point: 1e-5 relative."""
import numpy as np
import pytest
import torch

from conftest import golden
from faster_rcnn_pytorch_b200 import modules, ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5
CASES = [("voc", dict(seed=8300, N=20646, S=128, C=21)), ("coco", dict(seed=8301, N=37350, S=128, C=81, n_pos=128, n_neg=128)),
         ("small", dict(seed=8302, N=1440, S=128, C=21, n_pos=3, n_neg=253, frc_pos=5))]


def dev(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t.requires_grad_(True) if grad else t


@pytest.mark.parametrize("name,kw", CASES)
def test_region_loss_reference_goldens(name, kw):
    g = golden("loss")
    x = synth.loss_inputs(**kw)
    S = x["frc_cls"].shape[0]
    rc, rr = dev(x["rpn_cls"][None], True), dev(x["rpn_reg"][None], True)
    fc, fr = dev(x["frc_cls"][None], True), dev(x["frc_reg"][None], True)
    loss = ops.region_loss(rc, rr, dev(x["rpn_tcls"][None]), dev(x["rpn_treg"][None]), fc, fr, dev(x["frc_tcls"][None]),
                           dev(x["frc_treg"][None]))
    assert loss.shape == (1, 5)
    np.testing.assert_allclose(loss[0].detach().cpu().numpy(), g[f"{name}_loss"], rtol=RTOL)
    loss[0, 0].backward()                                   # total_loss.backward() like train.py
    valid = np.nonzero(x["rpn_tcls"] >= 0)[0]
    grc, grr = rc.grad[0].cpu().numpy(), rr.grad[0].cpu().numpy()
    np.testing.assert_allclose(grc[valid], g[f"{name}_g_rpn_cls_valid"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(grr[valid], g[f"{name}_g_rpn_reg_valid"], rtol=RTOL, atol=1e-9)
    assert not grc[x["rpn_tcls"] < 0].any() and not grr[x["rpn_tcls"] <= 0].any()
    np.testing.assert_allclose(fc.grad[0].cpu().numpy(), g[f"{name}_g_frc_cls"], rtol=RTOL, atol=1e-8)
    gfr = fr.grad[0].cpu().numpy()
    np.testing.assert_allclose(gfr[np.arange(S), x["frc_tcls"]], g[f"{name}_g_frc_reg_rows"], rtol=RTOL, atol=1e-9)
    assert np.count_nonzero(gfr) == np.count_nonzero(g[f"{name}_g_frc_reg_rows"])


def test_region_loss_modules_match_reference_call_shapes():
    """modules.FRCNNLoss(pred, target) with the tuples of FRCNN.forward: pred_reg already gathered to [128,4]."""
    g = golden("loss")
    x = synth.loss_inputs(seed=8300, N=20646, S=128, C=21)
    S = 128
    gathered = x["frc_reg"][np.arange(S), x["frc_tcls"]]
    pred = (dev(x["rpn_cls"][None], True), dev(x["rpn_reg"][None], True), dev(x["frc_cls"], True), dev(gathered, True))
    target = (dev(x["rpn_tcls"]), dev(x["rpn_treg"]), dev(x["frc_tcls"]), dev(x["frc_treg"]))
    out = modules.FRCNNLoss(None)(pred, target)
    np.testing.assert_allclose([float(v.detach()) for v in out], g["voc_loss"], rtol=RTOL)
    out[0].backward()
    np.testing.assert_allclose(pred[3].grad.cpu().numpy(), g["voc_g_frc_reg_rows"], rtol=RTOL, atol=1e-9)
    a, b = modules.RPNLoss()(pred[0].detach(), pred[1].detach(), target[0], target[1])
    c, d = modules.FastRCNNLoss()(pred[2].detach(), dev(x["frc_reg"].reshape(S, -1)), target[2], target[3])
    np.testing.assert_allclose([float(a), float(b), float(c), float(d)], g["voc_loss"][1:], rtol=RTOL)


def test_region_loss_batched_padding_and_upstream_scaling(oracle):
    """Batch of 3 images; image 2 has a short Fast R-CNN sample (padding rows = -1); per-component upstream gradients."""
    xs = [synth.loss_inputs(seed=8400 + i, N=5000, S=128, C=21, n_pos=10 + i, n_neg=200) for i in range(3)]
    xs[2]["frc_tcls"][90:] = -1
    st = lambda k, grad=False: dev(np.stack([x[k] for x in xs]), grad)
    rc, rr, fc, fr = st("rpn_cls", True), st("rpn_reg", True), st("frc_cls", True), st("frc_reg", True)
    loss = ops.region_loss(rc, rr, st("rpn_tcls"), st("rpn_treg"), fc, fr, st("frc_tcls"), st("frc_treg"))
    for i, x in enumerate(xs):
        want = oracle.region_loss(x["rpn_cls"], x["rpn_reg"], x["rpn_tcls"], x["rpn_treg"], x["frc_cls"], x["frc_reg"],
                                  x["frc_tcls"], x["frc_treg"])
        np.testing.assert_allclose(loss[i].detach().cpu().numpy(), want, rtol=RTOL)
    w = torch.tensor([[0.0, 2.0, 0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0, 3.0], [0.0, 0.0, 0.0, 1.0, 0.0]], device=DEV)
    (loss * w).sum().backward()
    assert not rr.grad[0].any() and rc.grad[0].any()            # image 0: only the RPN class loss was weighted
    assert not fr.grad[2].any() and not rc.grad[2].any() and fc.grad[2].any()
    assert not fc.grad[2, 90:].any()                            # padding rows get no gradient
    # torch autograd of the reference formulas on the GPU as the gradient checker for image 1 (total x1, frcnn_reg x3 more)
    x = xs[1]
    t = {k: dev(v) for k, v in x.items()}
    p = [t[k].clone().requires_grad_(True) for k in ("rpn_cls", "rpn_reg", "frc_cls", "frc_reg")]
    ce = torch.nn.functional.cross_entropy

    def sl1(a, b, beta):
        d = (a - b).abs()
        return torch.where(d >= beta, d - 0.5 * beta, 0.5 * d ** 2 / beta)
    nv = (t["rpn_tcls"] >= 0).sum()
    l1 = ce(p[0], t["rpn_tcls"], ignore_index=-1)
    l2 = sl1(p[1][t["rpn_tcls"] > 0], t["rpn_treg"][t["rpn_tcls"] > 0], 1 / 9).sum() / nv
    l3 = ce(p[2], t["frc_tcls"])
    rows = p[3][torch.arange(128), t["frc_tcls"]]
    l4 = sl1(rows[t["frc_tcls"] > 0], t["frc_treg"][t["frc_tcls"] > 0], 1.0).sum() / 128
    (l1 + l2 + l3 + 4.0 * l4).backward()
    for got, want in zip((rc.grad[1], rr.grad[1], fc.grad[1], fr.grad[1]), (q.grad for q in p)):
        np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=2e-5, atol=1e-8)


def test_region_loss_refuses_cpu_tensors():
    x = synth.loss_inputs(seed=1, N=100, S=8, C=5, n_pos=3, n_neg=5, frc_pos=2)
    with pytest.raises(ValueError):
        ops.region_loss(torch.from_numpy(x["rpn_cls"][None]), torch.from_numpy(x["rpn_reg"][None]),
                        torch.from_numpy(x["rpn_tcls"][None]), torch.from_numpy(x["rpn_treg"][None]))
