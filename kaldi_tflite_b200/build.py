"""
Builds libktf_b200.so in-tree (kaldi_tflite_b200/lib/) with nvcc for sm_100a.

    python -m kaldi_tflite_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIBNAME = "libktf_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_dep_mtime():
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    out = lib_path()
    dep_t = _newest_dep_mtime()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= dep_t:
        return out

    def compile_one(src):
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= dep_t):
            return obj, ""
        cmd = [NVCC] + ARCH_FLAGS + COMMON + ["-c", os.path.join(CSRC, src), "-o", obj]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}")
        return obj, p.stdout

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    if verbose:
        for _, log in results:
            if log:
                print(log)
    objs = [o for o, _ in results]
    cmd = [NVCC] + ARCH_FLAGS + ["-shared", "-o", out] + objs + ["-ldl"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"link failed:\n{p.stdout}")
    return out


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
