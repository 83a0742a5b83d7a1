"""Builds the C++ host layer: libkwage_host.so (for ctypes) and the CLIs kwage / kwage_tools,
all linked against libkwage_cuda.so (rpath'd, in-tree)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB_DIR = os.path.join(PKG, "lib")
BIN_DIR = os.path.join(PKG, "bin")
HOST_LIB = os.path.join(LIB_DIR, "libkwage_host.so")
COMMON = ["formats.cpp", "stages.cpp"]
HEADERS = [os.path.join(HERE, "kwage_host.h"), os.path.join(os.path.dirname(PKG), "include", "kwage_cuda.h")]


def _cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _stale(out, deps):
    return (not os.path.exists(out)) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps)


def build(force=False):
    from kwage_b200 import build as kbuild
    kbuild.build()
    os.makedirs(BIN_DIR, exist_ok=True)
    flags = ["-O2", "-std=c++11", "-Wall", "-fPIC", "-pthread"]
    link = ["-L" + LIB_DIR, "-lkwage_cuda", "-lz", "-Wl,-rpath," + LIB_DIR, "-Wl,-rpath,$ORIGIN/../lib", "-Wl,-rpath,$ORIGIN"]
    common = [os.path.join(HERE, s) for s in COMMON]
    targets = [
        (HOST_LIB, common + [os.path.join(HERE, "host_capi.cpp")], ["-shared"]),
        (os.path.join(BIN_DIR, "kwage"), common + [os.path.join(HERE, "kwage_main.cpp")], []),
        (os.path.join(BIN_DIR, "kwage_tools"), common + [os.path.join(HERE, "tools_main.cpp")], []),
    ]
    for out, srcs, extra in targets:
        if force or _stale(out, srcs + HEADERS + [kbuild.LIB_PATH]):
            subprocess.run([_cxx()] + flags + extra + ["-o", out] + srcs + link, check=True)
    return HOST_LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(HOST_LIB)
