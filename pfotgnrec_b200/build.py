"""Build libpfo_b200.so in-tree with nvcc for sm_100a (no torch linkage: plain C ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpfo_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
# per-file extra flags: the MV kernel must not contract a*b+c into fma (fp64 parity with numpy)
SOURCES = {
    "graph_kernels.cu": [],
    "linear_simt.cu": [],
    "linear_tc.cu": [],
    "linear_tma.cu": [],
    "wgrad_tma.cu": [],
    "memory_kernels.cu": [],
    "route_kernels.cu": [],
    "peer_kernels.cu": [],
    "attention_kernels.cu": [],
    "fold_kernels.cu": [],
    "mv_kernels.cu": ["-fmad=false"],
    "eval_kernels.cu": ["-fmad=false"],     # fp64 metric block rounds like numpy
}


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "pfo_b200.h"))
    objs = []
    for src, extra in SOURCES.items():
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [sp] + headers):
            cmd = [nvcc, "-c", sp, "-o", obj] + ARCH + COMMON + extra
            if verbose:
                cmd += ["-Xptxas", "-v"]
            print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-Xcompiler", "-fPIC"]
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
