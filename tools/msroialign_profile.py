"""Developer tool: MultiScaleRoIAlign at FPN shapes (800x1333 input, strides 4..32, C = 256, 512 rois per image) next to
torchvision's CUDA implementation; forward and forward + backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torchvision
from faster_rcnn_pytorch_b200 import ops
dev = torch.device("cuda:0")
B, C, hw, per = 2, 256, (800, 1344), 512
rs = np.random.RandomState(5)
feats = [torch.from_numpy(rs.standard_normal((B, C, hw[0] // s, hw[1] // s)).astype(np.float32)).to(dev) for s in (4, 8, 16, 32)]
side = np.exp(rs.uniform(np.log(16), np.log(600), B * per)); ar = np.exp(rs.uniform(-0.7, 0.7, B * per))
w, h = side * np.sqrt(ar), side / np.sqrt(ar)
cx, cy = rs.uniform(0, hw[1], B * per), rs.uniform(0, hw[0], B * per)
x1, y1 = np.clip(cx - w / 2, 0, hw[1] - 2), np.clip(cy - h / 2, 0, hw[0] - 2)
x2, y2 = np.clip(cx + w / 2, x1 + 1, hw[1]), np.clip(cy + h / 2, y1 + 1, hw[0])
boxes = [torch.from_numpy(np.stack([x1, y1, x2, y2], 1)[b * per:(b + 1) * per].astype(np.float32)).to(dev) for b in range(B)]
shapes = [hw] * B
tv = torchvision.ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2)
x = {str(i): f for i, f in enumerate(feats)}

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

a = ops.multiscale_roi_align(feats, boxes, shapes, 7, 2)
bq = tv(x, boxes, shapes)
print("max abs diff vs torchvision CUDA", float((a - bq).abs().max()))
print("fwd us: ours", round(t(lambda: ops.multiscale_roi_align(feats, boxes, shapes, 7, 2)), 1),
      "torchvision", round(t(lambda: tv(x, boxes, shapes)), 1))
fg = [f.clone().requires_grad_(True) for f in feats]
xg = {str(i): f for i, f in enumerate(fg)}
go = torch.randn_like(a)
def ours_fb():
    ops.multiscale_roi_align(fg, boxes, shapes, 7, 2).backward(go)
def tv_fb():
    tv(xg, boxes, shapes).backward(go)
print("fwd+bwd us: ours", round(t(ours_fb), 1), "torchvision", round(t(tv_fb), 1))
