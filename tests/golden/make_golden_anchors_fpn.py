"""Golden fixtures for the FPN variant's anchors (run in the BUILD container only).

    python tests/golden/make_golden_anchors_fpn.py

Runs the reference's own two anchor lines (models/new_model.py:23-25,43-44: torchvision's AnchorGenerator on an ImageList,
then the division by (w, h, w, h)) for several image sizes and stores, per size, the sha256 of the [N,4] fp32 array, N,
and its first / last rows (tests/golden/anchors_fpn.npz); plus the per-level base anchors.
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torchvision
from torchvision.models.detection.image_list import ImageList

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def main():
    gen = torchvision.models.detection.rpn.AnchorGenerator(sizes=((32,), (64,), (128,), (256,), (512,)),
                                                           aspect_ratios=((0.5, 1.0, 2.0),) * 5)
    g = {"base": np.stack([b.numpy() for b in gen.cell_anchors])}
    for h, w in [(128, 192), (160, 160), (800, 1344), (800, 1333), (600, 1000), (97, 131)]:
        x = torch.zeros(1, 3, h, w)
        feats = [torch.zeros(1, 1, -(-h // s), -(-w // s)) for s in (4, 8, 16, 32, 64)]
        anchor = gen(ImageList(x, [(w, h)]), feats)[0]
        anchor = anchor / torch.FloatTensor([w, h, w, h])          # models/new_model.py:44
        a = anchor.numpy()
        key = f"{h}x{w}"
        g[f"sha_{key}"] = sha(a)
        g[f"n_{key}"] = np.int64(a.shape[0])
        g[f"head_{key}"] = a[:6].copy()
        g[f"tail_{key}"] = a[-6:].copy()
    np.savez_compressed(os.path.join(HERE, "anchors_fpn.npz"), **g)
    print("anchors_fpn.npz", os.path.getsize(os.path.join(HERE, "anchors_fpn.npz")))


if __name__ == "__main__":
    main()
