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
SOURCES = ["tables.cu", "generic.cu", "fast_fwd.cu", "fast_400.cu", "fast_512.cu", "fast_1024.cu", "fast_800.cu", "fast_inv.cu", "small.cu", "resample.cu", "post.cu", "api.cu"]
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
    # straight-line mel code for the named filterbanks, generated from the product's own b2a_mel_filters
    gen = os.path.join(HERE, "mel_gen.cuh")
    if force or not os.path.exists(gen) or os.path.getmtime(gen) < max(
            os.path.getmtime(os.path.join(HERE, "tables.cu")), os.path.getmtime(os.path.join(HERE, "gen_mel.py"))):
        sys.path.insert(0, HERE)
        try:
            import gen_mel
            gen_mel.generate(gen)
        finally:
            sys.path.pop(0)
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    cflags = [f for f in FLAGS if f not in ("--use_fast_math=false", "-shared")]
    if verbose:
        cflags += ["-Xptxas", "-v"]
    if os.environ.get("B2A_DEV_400_ONLY"):
        cflags += ["-DB2A_DEV_400_ONLY"]
    if os.environ.get("B2A_PHASE_CLOCKS"):
        cflags += ["-DB2A_PHASE_CLOCKS"]
    # development: A/B builds — extra -D flags and a separate output library (selected at run time with B2A_LIB=path)
    cflags += os.environ.get("B2A_BUILD_DEFS", "").split()
    out_lib = os.environ.get("B2A_BUILD_OUT", LIB)
    # one nvcc per translation unit, in parallel; then one link
    obj_dir = os.path.join(HERE, "_obj") if out_lib == LIB else os.path.join(HERE, "_obj", os.path.basename(out_lib))
    os.makedirs(obj_dir, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        return obj, subprocess.run([nvcc()] + cflags + ["-c", "-o", obj, src], capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        results = list(ex.map(compile_one, srcs))
    log = ""
    for obj, r in results:
        log += r.stdout + r.stderr
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building libb200audio.so")
    r = subprocess.run([nvcc(), "-shared", "-cudart", "static", "-o", out_lib] + [o for o, _ in results],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libb200audio.so")
    r.stderr = log + r.stderr
    if verbose:
        print(r.stderr)
    return out_lib


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
