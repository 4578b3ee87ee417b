"""Build recipe for libmsweep_b200.so (the C-ABI CUDA library) and the C++ host tools.

Everything is compiled IN-TREE for sm_100a with nvcc; nothing is JIT-compiled at import time and
there is no fallback: if the library is missing the package raises.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIBDIR = os.path.join(HERE, "lib")
BINDIR = os.path.join(HERE, "bin")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libmsweep_b200.so")
CLI = os.path.join(BINDIR, "mSWEEP_b200")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
CU_SOURCES = ["ctx.cu", "ec_build.cu", "likelihood.cu", "vi.cu", "vi_batch.cu", "bootstrap.cu", "mt64_jump.cu", "assign.cu"]
HEADERS = ["common.cuh", "peer.cuh", "handles.cuh", "vi_kernels.cuh", "vi_sparse_rcg.cuh", "mathfn.cuh", os.path.join("..", "..", "include", "msweep_b200.h")]


def _mtime(path: str) -> float:
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def _stale(target: str, deps: list[str]) -> bool:
    t = _mtime(target)
    return t == 0.0 or any(_mtime(d) > t for d in deps)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + os.path.basename(cmd[-1]))


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    objs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append(NVCC.split() + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(_run, jobs))
    if force or jobs or _stale(LIB, objs):
        # cudart is linked statically (nvcc default); NCCL is dlopen'ed at run time (ctx.cu)
        _run(NVCC.split() + ["-shared", "-ccbin", HOST_CXX, "-o", LIB] + objs + ["-ldl", "-lpthread"])
    return LIB


def build_host(force: bool = False) -> str | None:
    """The C++17 host side: mSWEEP-compatible driver over the C ABI (host/*.cpp)."""
    main = os.path.join(HOST, "main.cpp")
    if not os.path.exists(main):
        return None
    os.makedirs(BINDIR, exist_ok=True)
    srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".cpp")]
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hpp")] + [LIB]
    if force or _stale(CLI, deps):
        _run([HOST_CXX, "-std=c++17", "-O2", "-fopenmp", "-Wall", "-I", os.path.join(HERE, "..", "include"), "-o", CLI] + srcs +
             ["-L", LIBDIR, "-lmsweep_b200", "-Wl,-rpath,$ORIGIN/../lib", "-lpthread", "-ldl", "-lz"])
    return CLI


def build_all(force: bool = False) -> None:
    build_library(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print(LIB)
