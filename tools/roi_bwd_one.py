"""Developer tool: one RoIPool forward + backward at the config-3 shape with RPN-sized rois (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from faster_rcnn_pytorch_b200 import ops, synth
dev = torch.device("cuda:0")
B, C, fh, fw, per = 16, 512, 37, 62, 128
K = B * per
feat = torch.from_numpy(synth.features(1, B, C, fh, fw)).to(dev)
go = torch.randn((K, C, 7, 7), device=dev)
rs = np.random.RandomState(5)
lo, hi = float(os.environ.get("ROI_LO", "8")), float(os.environ.get("ROI_HI", "30"))
w = rs.uniform(lo, hi, K); h = rs.uniform(lo, min(hi, fh - 1), K)
x1 = rs.uniform(0, fw - w); y1 = rs.uniform(0, fh - h)
r = torch.from_numpy(np.stack([np.repeat(np.arange(B), per), x1, y1, x1 + w, y1 + h], 1).astype(np.float32)).to(dev)
for _ in range(2):
    out, arg = ops.roi_pool_forward(feat, r)
    gin = ops.roi_pool_backward(go, arg, r, feat.shape)
torch.cuda.synchronize()
print("ok")
