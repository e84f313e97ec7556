"""N > 1 host logic on CPU (gloo, world_size 2 and 3): image sharding and the fixed-shape detection all-gather
that replaces the reference's pickled all_gather (util/misc.py:89-129)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from faster_rcnn_pytorch_b200 import dist as fdist


def test_shard_range_partitions_every_image_once():
    for n in (0, 1, 7, 8, 64, 65, 1001):
        for ws in (1, 2, 3, 4, 8):
            seen = []
            for r in range(ws):
                lo, hi = fdist.shard_range(n, r, ws)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [fdist.shard_range(n, r, ws)[1] - fdist.shard_range(n, r, ws)[0] for r in range(ws)]
            assert max(sizes) - min(sizes) <= 1


def _fake_detections(image_id: int, cap: int = 12):
    rs = np.random.RandomState(1000 + image_id)
    n = int(rs.randint(0, cap + 1))
    boxes = np.zeros((cap, 4), np.float32); labels = np.full((cap,), -1, np.int32); scores = np.zeros((cap,), np.float32)
    boxes[:n] = rs.uniform(0, 1, (n, 4)).astype(np.float32)
    labels[:n] = np.sort(rs.randint(0, 20, n)).astype(np.int32)          # class-major like _suppress
    scores[:n] = rs.uniform(0.05, 1, n).astype(np.float32)
    return boxes, labels, scores, n


def _pack_np(db, dl, ds, dc, max_det):
    """Test-side restatement of the packing (the product packs with a CUDA kernel, checked in test_gpu_post.py)."""
    B, cap = dl.shape
    out = torch.zeros((B, max_det, 6), dtype=torch.float32)
    cnt = dc.clamp(max=min(max_det, cap)).to(torch.int32)
    for i in range(B):
        k = int(cnt[i])
        out[i, :k, :4] = db[i, :k]
        out[i, :k, 4] = ds[i, :k]
        out[i, :k, 5] = dl[i, :k].to(torch.float32)
    return out, cnt


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, n_images, max_det, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        lo, hi = fdist.shard_range(n_images)
        imgs = [_fake_detections(i) for i in range(lo, hi)]
        cap = 12
        if imgs:
            db = torch.from_numpy(np.stack([x[0] for x in imgs])); dl = torch.from_numpy(np.stack([x[1] for x in imgs]))
            ds = torch.from_numpy(np.stack([x[2] for x in imgs])); dc = torch.tensor([x[3] for x in imgs], dtype=torch.int32)
        else:
            db = torch.zeros((0, cap, 4)); dl = torch.zeros((0, cap), dtype=torch.int32); ds = torch.zeros((0, cap))
            dc = torch.zeros((0,), dtype=torch.int32)
        packed, cnt = _pack_np(db, dl, ds, dc, max_det)
        ids = torch.arange(lo, hi, dtype=torch.int64)
        P, C, I = fdist.gather_detections(packed, cnt, ids)
        if n_images % ws == 0:      # equal shards: the sync-free path must give the same result
            P2, C2, I2 = fdist.gather_detections(packed, cnt, ids, equal_batch=True)
            assert torch.equal(P, P2) and torch.equal(C, C2) and torch.equal(I, I2)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), P=P.numpy(), C=C.numpy(), I=I.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws,n_images", [(2, 9), (2, 8), (3, 7), (2, 1)])
def test_gather_detections_matches_single_process(tmp_path, ws, n_images):
    max_det = 10
    mp.spawn(_worker, args=(ws, _free_port(), n_images, max_det, str(tmp_path)), nprocs=ws, join=True)
    # single-process expectation: every image once, in image order (shards are contiguous and rank ordered)
    imgs = [_fake_detections(i) for i in range(n_images)]
    db = torch.from_numpy(np.stack([x[0] for x in imgs])); dl = torch.from_numpy(np.stack([x[1] for x in imgs]))
    ds = torch.from_numpy(np.stack([x[2] for x in imgs])); dc = torch.tensor([x[3] for x in imgs], dtype=torch.int32)
    want_p, want_c = _pack_np(db, dl, ds, dc, max_det)
    for r in range(ws):
        g = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert np.array_equal(g["I"], np.arange(n_images))
        assert np.array_equal(g["C"], want_c.numpy())
        assert np.array_equal(g["P"], want_p.numpy())
