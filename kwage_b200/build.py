"""Builds libkwage_cuda.so (sm_100a) in-tree with nvcc.  No torch dependency: the product library is
plain CUDA runtime behind a C ABI (include/kwage_cuda.h)."""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libkwage_cuda.so")
SOURCES = ["api.cu", "bloom_build.cu", "transpose.cu", "search.cu", "synth.cu", "crc32.cu", "merge.cu"]
HEADERS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + \
    [os.path.join(ROOT, "include", "kwage_cuda.h"), os.path.abspath(__file__)]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "--use_fast_math",        # integer kernels; only affects the two float ops we pin with __fmul_rn
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, ptxas_info=False):
    # KWG_NVCC_EXTRA: extra nvcc flags for A/B builds of experiment macros (profiles/run/*.sh); forces a rebuild
    extra = os.environ.get("KWG_NVCC_EXTRA", "").split()
    force = force or bool(extra)
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(PKG, "build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = nvcc_path()
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(obj_dir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + HEADERS):
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or ptxas_info or verbose:
            sys.stderr.write(out)
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    print(LIB_PATH)
