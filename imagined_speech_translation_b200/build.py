"""Builds libeegx.so in-tree with nvcc for sm_100a (no JIT cache, no torch arch list).

    python -m imagined_speech_translation_b200.build [--force] [--verbose]

Every csrc/*.cu is compiled to its own object (in parallel, only when it or a header changed) and the
objects are linked into one shared library.  The .so lands next to this file so it travels to the
GPU box with the snapshot; the objects live under build/ (git-ignored, gpurun-ignored).
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libeegx.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) \
        + glob.glob(os.path.join(ROOT, "include", "*.h")) + [os.path.abspath(__file__)]


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdrs = _headers()
    srcs = sources()
    todo = [s for s in srcs if force or _stale(_obj(s), [s] + hdrs)]
    if not todo and not _stale(LIB, [_obj(s) for s in srcs]):
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", _obj(src)]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    jobs = max(1, min(len(todo), int(os.environ.get("EEGX_BUILD_JOBS", os.cpu_count() or 4))))
    failed = False
    with ThreadPoolExecutor(jobs) as pool:
        for src, res in pool.map(compile_one, todo):
            if verbose or res.returncode != 0:
                sys.stderr.write(f"== {os.path.basename(src)}\n{res.stdout}{res.stderr}")
            failed |= res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libeegx.so (see stderr)")
    link = [nvcc, "-shared", "-o", LIB] + [_obj(s) for s in srcs] + ["-lcuda"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("linking libeegx.so failed (see stderr)")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
