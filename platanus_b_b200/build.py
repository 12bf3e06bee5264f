"""Build libpbk.so in-tree (nvcc, sm_100a only).  Usable as `python -m platanus_b_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PBK_LIB_TAG / PBK_EXTRA_CFLAGS: tuning builds only (scripts/tune_variants.sh) -- a second library next to the product one
_TAG = os.environ.get("PBK_LIB_TAG", "")
OUT_DIR = os.path.join(HERE, "_lib" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(OUT_DIR, "libpbk.so")

NVCC = os.environ.get("PBK_NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = (["-DPBK_EXPERIMENT"] if os.environ.get("PBK_EXPERIMENT") else []) + os.environ.get("PBK_EXTRA_CFLAGS", "").split() + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"]
UNITS = ["pbk_kernels.cu", "pbk_api.cu", "pbk_group.cu", "pbk_host.cpp"]
HEADERS = ["pbk_device.cuh", "pbk_kernels.cuh", "pbk_kernels_impl.cuh", os.path.join("..", "..", "include", "pbk.h")]


def _stale(target: str, deps: list) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    env = dict(os.environ)
    env.pop("CXX", None)        # the image exports a wrapper CXX that breaks host linking
    env.pop("CC", None)

    def compile_one(unit: str) -> str:
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OUT_DIR, os.path.splitext(unit)[0] + ".o")
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC, *ARCH, *CFLAGS, "-ccbin", "g++", "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), file=sys.stderr)
            subprocess.run(cmd, check=True, env=env)
        return obj

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(compile_one, UNITS))
    if force or _stale(LIB, objs):
        subprocess.run([NVCC, *ARCH, "-ccbin", "g++", "-shared", "-o", LIB, *objs, "-cudart", "static", "-lpthread"],
                       check=True, env=env)
    return LIB


CLI = os.path.join(OUT_DIR, "pbk_assemble")
HOST = os.path.join(HERE, "host")


def build_cli(force: bool = False) -> str:
    """The host C++ side of the drop-in (host/pbk_counter.hpp + host/pbk_assemble.cpp), linked against libpbk.so."""
    lib = build_lib()
    srcs = [os.path.join(HOST, "pbk_assemble.cpp"), os.path.join(HOST, "pbk_counter.hpp"),
            os.path.join(HERE, "..", "include", "pbk.h")]
    if force or _stale(CLI, srcs + [lib]):
        env = dict(os.environ)
        env.pop("CXX", None)
        env.pop("CC", None)
        subprocess.run(["g++", "-O2", "-std=c++11", "-Wall", "-Wextra", "-pthread", "-o", CLI, srcs[0],
                        "-L" + OUT_DIR, "-lpbk", "-Wl,-rpath,$ORIGIN"], check=True, env=env)
    return CLI


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_cli(force="--force" in sys.argv))
