"""Builds libfrr.so in-tree with nvcc for sm_100a (the built .so travels to the GPU box).

    python -m faster_rcnn_pytorch_b200.build [--force] [--verbose]

Every translation unit is compiled on its own (in parallel, only when it or a header changed) into
``csrc/_obj/`` and the objects are linked into ``libfrr.so``.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(PKG, "libfrr.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # parity: no FMA contraction anywhere (hot ops use *_rn intrinsics anyway)
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "-I", os.path.join(REPO, "include"),
    "-I", CSRC,
]
# the CUDA runtime comes from the libcudart.so.12 that torch has already loaded into the process (same major
# version; see _lib.load) -- the library carries no second, statically linked runtime
LINK_FLAGS = ["--shared", "-cudart", "shared"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(REPO, "include", "frr.h")]


def _obj_of(src: str) -> str:
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return _stale(LIB, sources() + _headers() + [os.path.abspath(__file__)])


def _run(cmd, env, verbose):
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd[-3:]))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers() + [os.path.abspath(__file__)]
    todo = [s for s in sources() if force or _stale(_obj_of(s), [s] + hdrs)]
    extra = ["-Xptxas", "-v"] if verbose else []
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(lambda s: _run([nvcc] + NVCC_FLAGS + extra + ["-c", "-o", _obj_of(s), s], env, verbose), todo))
    live = {_obj_of(s) for s in sources()}
    for o in glob.glob(os.path.join(OBJ, "*.o")):
        if o not in live:
            os.unlink(o)
    _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a"] + LINK_FLAGS + ["-o", LIB] + sorted(live), env, verbose)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
