"""CPU: the model of the bucketed NMS kernel's algorithm (tests/nms_model.py: chunks of 2048 candidates, pair tests
restricted to admissible (area class, x bin) keys, screen before exact test, predecessor fix-point) keeps exactly
what the oracle (= torchvision's CPU kernel) keeps, on RPN-like boxes and on sets built to sit on the pruning
boundaries.  This pins the exactness of the pruning rules independently of the CUDA transcription."""
import numpy as np
import pytest

import nms_cases
import nms_model


@pytest.mark.parametrize("thr", [0.7, 0.5, 0.3])
@pytest.mark.parametrize("name", nms_cases.ALL)
def test_bucketed_model_equals_oracle(oracle, name, thr):
    n = 3000 if name != "rpn_like" else 5000
    b = nms_cases.make(name, 11, n, thr)
    want = oracle.nms(b, -np.arange(len(b), dtype=np.float32), thr)[:1000]
    for chunks in [(2048,), (512, 1024, 256)]:
        got = nms_model.nms_bucketed(b, thr, 1000, chunk_sizes=chunks)
        assert np.array_equal(got, want), (name, thr, chunks)


def test_model_constants_match_kat_thresholds():
    t = nms_model.make_thr(0.7)
    assert float(t["up"]) > 0.7 and float(np.nextafter(t["up"], np.float32(0))) <= 0.7
    t3 = nms_model.make_thr(0.3)
    assert float(t3["up"]) == float(np.float32(0.3))        # 0.3f > 0.3: IoU == 0.3f is suppressed (KAT-1)
