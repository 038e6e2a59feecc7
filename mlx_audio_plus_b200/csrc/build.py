"""Builds libb200audio.so in-tree with nvcc for sm_100a (no torch involved: the library is a plain
CUDA-runtime C-ABI shared object).  Used by __graft_entry__.build() and `python -m mlx_audio_plus_b200.csrc.build`."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libb200audio.so")
SOURCES = ["tables.cu", "generic.cu", "fast_fwd.cu", "small.cu", "api.cu"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared", "-cudart", "static",
]


def nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(PKG), "include", "b200audio.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    flags = [f for f in FLAGS if f != "--use_fast_math=false"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    cmd = [nvcc()] + flags + ["-o", LIB] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libb200audio.so")
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
