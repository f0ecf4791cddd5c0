"""In-tree build of libzkb.so (hand-written sm_100a CUDA + the C ABI of include/zkb.h).

nvcc cross-compiles without a GPU.  Objects land in eigen-zeth_b200/build/ (git-ignored), the
library in eigen-zeth_b200/libzkb.so (git-ignored, but it travels with the gpurun snapshot).
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libzkb.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
                  "--expt-relaxed-constexpr"]
CXXFLAGS = ["-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-fvisibility=hidden"]


def _sources():
    out = []
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(".cu") or f.endswith(".cpp"):
            out.append(os.path.join(CSRC, f))
    return out


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps_mtime(path, seen=None):
    """newest mtime of `path` and of everything it #includes with quotes (recursively)."""
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return 0.0
    seen.add(path)
    m = os.path.getmtime(path)
    with open(path) as f:
        for inc in _INC.findall(f.read()):
            m = max(m, _deps_mtime(os.path.join(os.path.dirname(path), inc), seen))
    return m


PER_FILE = {}     # per-file extra flags (none at present)


def _compile(src, obj):
    if src.endswith(".cu"):
        cmd = [NVCC] + CUFLAGS + PER_FILE.get(os.path.basename(src), []) + ["-c", src, "-o", obj]
    else:
        cmd = ["g++"] + CXXFLAGS + ["-I/usr/local/cuda/include", "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("compile failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(BUILD, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < _deps_mtime(src):
            jobs.append((src, obj))
    if jobs:
        if verbose:
            print("compiling", [os.path.basename(s) for s, _ in jobs], file=sys.stderr)
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _compile(*j), jobs))
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xlinker", "--exclude-libs,ALL"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
