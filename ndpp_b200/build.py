"""Builds libndppgpu.so (hand-written sm_100a CUDA + the C-ABI host layer) in-tree with nvcc.

-fmad=false: the reference is compiled for x86-64 without FMA contraction and several of its closed
forms cancel catastrophically (csrc/legendre.cuh); contracting a*b+c on the device would move the
results outside the parity tolerance.  Explicit fma() calls (the FP64 peak micro-benchmark) are kept.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(CSRC, "libndppgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
         "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++"]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".inc"))] + \
        [os.path.join(os.path.dirname(HERE), "include", "ndppgpu.h")]


def build(force: bool = False, verbose: bool = False, extra=()) -> str:
    srcs = sources()
    if not force and os.path.exists(SO) and all(os.path.getmtime(s) <= os.path.getmtime(SO) for s in srcs):
        return SO
    cmd = [NVCC] + FLAGS + list(extra) + os.environ.get("NDPP_NVCC_EXTRA", "").split() + \
        ["-o", SO, os.path.join(CSRC, "ndppgpu.cu")]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return SO


TOOL_SRC = os.path.join(os.path.dirname(HERE), "tools", "ndpp_calc_scatt.cpp")
TOOL = os.path.join(os.path.dirname(HERE), "tools", "ndpp_calc_scatt")


def build_tool(force: bool = False) -> str:
    """The C++ driver above the C-ABI (tools/ndpp_calc_scatt.cpp over include/ndpp_host.hpp), linked against the
    in-tree libndppgpu.so with an $ORIGIN-relative rpath so that the pair travels to the GPU box."""
    inc = os.path.join(os.path.dirname(HERE), "include")
    deps = [TOOL_SRC, os.path.join(inc, "ndpp_host.hpp"), os.path.join(inc, "ndpp_library.hpp"), os.path.join(inc, "ndppgpu.h"), SO]
    if not force and os.path.exists(TOOL) and all(os.path.getmtime(d) <= os.path.getmtime(TOOL) for d in deps):
        return TOOL
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-I", inc, TOOL_SRC, "-o", TOOL, "-L", CSRC,
           "-lndppgpu", "-Wl,-rpath,$ORIGIN/../ndpp_b200/csrc"]
    subprocess.check_call(cmd)
    return TOOL


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else [])
    build_tool(force=True)
