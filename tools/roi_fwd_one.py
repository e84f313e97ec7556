import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from faster_rcnn_pytorch_b200 import ops, synth
dev = torch.device("cuda:0")
B, C, fh, fw, per = 16, 512, 37, 62, 128
feat = torch.from_numpy(synth.features(1, B, C, fh, fw)).to(dev)
rois = torch.from_numpy(np.concatenate([np.concatenate([np.full((per, 1), b, np.float32), synth.random_rois(10 + b, per, fh, fw, 1)[:, 1:]], 1) for b in range(B)])).to(dev)
for _ in range(2):
    out, arg = ops.roi_pool_forward(feat, rois)
torch.cuda.synchronize()
print("ok")
