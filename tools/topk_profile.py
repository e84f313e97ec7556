"""Developer tool: top-k kernel timing + per-phase cycle breakdown on the config-2 workload (GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from faster_rcnn_pytorch_b200 import ops, synth, _lib

HW = (608, 1008)
dev = torch.device("cuda:0")
B = int(os.environ.get("TOPK_B", "64"))
ins = [synth.rpn_head_outputs(2000 + i, HW) for i in range(B)]
reg = torch.from_numpy(np.stack([x[1] for x in ins])).to(dev)
sc = torch.from_numpy(np.stack([x[2] for x in ins])).to(dev)
boxes, scores, valid = ops.rpn_decode(reg, sc, image_hw=HW)
lib = _lib.load()
N = scores.shape[1]
names = ["r_load", "r_select", "r_compact", "r_sort", "r_write", "-", "-", "-", "load", "hist", "scan", "scatter", "sort", "write"]
for k in (12000, 6000):
    oi = torch.empty((B, k), dtype=torch.int32, device=dev)
    ob = None if os.environ.get("TOPK_NOBOX") else torch.empty((B, k, 4), dtype=torch.float32, device=dev)  # fused path: indices only
    oc = torch.empty((B,), dtype=torch.int32, device=dev)
    d = torch.zeros(16, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def run(dbg):
        _lib.check(lib.frr_topk_desc_profile(scores.data_ptr(), valid.data_ptr(), boxes.data_ptr(), B, N, k, None, oi.data_ptr(),
                                             None, ob.data_ptr() if ob is not None else None, oc.data_ptr(), dbg, st), "topk")
    for _ in range(3):
        run(d.data_ptr())
    d.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run(d.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    out = {"k": k, "B": B, "us": round(e0.elapsed_time(e1) / 20 * 1e3, 1)}
    out.update({n: int(v) // 20 for n, v in zip(names, d.cpu().tolist()) if n != "-"})
    print(json.dumps(out), flush=True)
