"""Golden vectors for the loss side: the UNMODIFIED reference losses/loss.py (RPNLoss, FastRCNNLoss, FRCNNLoss) with the
class-row gather of models/model.py:340-341, run on CPU with torch autograd -> tests/golden/loss.npz.

    python tests/golden/make_golden_loss.py        (needs /root/reference; the GPU box only reads the .npz)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")
from faster_rcnn_pytorch_b200 import synth  # noqa: E402
from losses.loss import FRCNNLoss  # noqa: E402  (the reference)

CASES = {"voc": dict(seed=8300, N=20646, S=128, C=21), "coco": dict(seed=8301, N=37350, S=128, C=81, n_pos=128, n_neg=128),
         "small": dict(seed=8302, N=1440, S=128, C=21, n_pos=3, n_neg=253, frc_pos=5)}


def main():
    g = {}
    crit = FRCNNLoss(None)
    for name, kw in CASES.items():
        x = synth.loss_inputs(**kw)
        S = x["frc_cls"].shape[0]
        rc = torch.from_numpy(x["rpn_cls"])[None].requires_grad_(True)
        rr = torch.from_numpy(x["rpn_reg"])[None].requires_grad_(True)
        fc = torch.from_numpy(x["frc_cls"]).requires_grad_(True)
        fr = torch.from_numpy(x["frc_reg"]).requires_grad_(True)
        tcls = torch.from_numpy(x["frc_tcls"])
        gathered = fr[torch.arange(0, S).long(), tcls.long()]                      # models/model.py:340-341
        out = crit((rc, rr, fc, gathered), (torch.from_numpy(x["rpn_tcls"]), torch.from_numpy(x["rpn_treg"]), tcls,
                                            torch.from_numpy(x["frc_treg"])))
        out[0].backward()
        g[f"{name}_loss"] = np.asarray([float(v) for v in out], np.float32)
        valid = np.nonzero(x["rpn_tcls"] >= 0)[0]
        assert not rc.grad[0].numpy()[x["rpn_tcls"] < 0].any() and not rr.grad[0].numpy()[x["rpn_tcls"] <= 0].any()
        g[f"{name}_g_rpn_cls_valid"] = rc.grad[0].numpy()[valid]
        g[f"{name}_g_rpn_reg_valid"] = rr.grad[0].numpy()[valid]
        g[f"{name}_g_frc_cls"] = fc.grad.numpy()
        g[f"{name}_g_frc_reg_rows"] = fr.grad.numpy()[np.arange(S), x["frc_tcls"]]
        assert np.count_nonzero(fr.grad.numpy()) == np.count_nonzero(g[f"{name}_g_frc_reg_rows"])
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **g)
    print("loss.npz", os.path.getsize(os.path.join(HERE, "loss.npz")), {k: g[k] for k in g if k.endswith("_loss")})


if __name__ == "__main__":
    main()
