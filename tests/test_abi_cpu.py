"""CPU-side checks of the drop-in boundary: libfrr.so loads without a GPU and exports exactly the
entry points include/frr.h declares; the product refuses CPU tensors (no fallback)."""
import ctypes

import numpy as np
import pytest
import torch

from faster_rcnn_pytorch_b200 import _lib, ops


def test_library_builds_loads_and_exports_every_declared_symbol():
    from faster_rcnn_pytorch_b200 import build
    build.build()
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, f"{n} declared in frr.h but not bound in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.frr_abi_version() == 1


def test_anchor_base_host_matches_oracle(oracle):
    assert np.array_equal(ops.anchor_base_table(), oracle.anchor_base_table())


def test_error_reporting_without_gpu():
    lib = _lib.load()
    rc = lib.frr_anchor_base_host(None, 16)
    assert rc == -1 and b"bad arguments" in lib.frr_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "frr_anchor_base_host")


def test_ops_refuse_cpu_tensors():
    with pytest.raises(ValueError, match="no CPU path"):
        ops.nms_sorted(torch.zeros(1, 4, 4), 0.5)
    with pytest.raises(ValueError, match="no CPU path"):
        ops.topk_desc(torch.zeros(1, 4), 2)


def test_product_does_not_import_the_oracle():
    import os, re
    pkg = os.path.dirname(_lib.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
