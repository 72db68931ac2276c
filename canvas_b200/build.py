"""Build libcanvasgpu.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The shared object lands in canvas_b200/_build/ (git-ignored, but it travels to the GPU box with the
repo snapshot).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libcanvasgpu.so")
SOURCES = ["capi.cu", "clean.cu", "wavelet.cu", "bin.cu", "cbs.cu", "hmm.cu", "merge.cu", "smooth.cu", "codec.cu", "normalize.cu", "bin_gc.cu", "comm.cu", "pedigree.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # the reference computes in IEEE double/float without fused multiply-add (.NET RyuJIT);
    # contraction would change results in the last bit
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"),
] + (["-DSEL_DEBUG"] if os.environ.get("CANVAS_SEL_DEBUG") else [])


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "canvasgpu.h"))
    objs, cmds = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmds.append([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    if cmds:  # one nvcc per translation unit, side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 1)) as ex:
            list(ex.map(subprocess.check_call, cmds))
    if force or _stale(LIB, objs):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
