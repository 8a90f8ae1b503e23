"""In-tree build of the native pieces (explicit nvcc / g++ commands, no build system).

    python -m fringe_b200.build            # libfringe_b200.so (+ host driver and bindings)

Outputs land next to the sources under fringe_b200/lib/ so the GPU box snapshot carries them.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
NVCC = os.environ.get("FRINGE_NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOST_CXX = os.environ.get("FRINGE_CXX", "/usr/bin/g++")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math=false", "-Xcompiler", "-fPIC",
              "-ccbin", HOST_CXX] + ARCH

CUDA_SOURCES = ["capi.cu", "nmap_kernels.cu", "evd_kernels.cu", "evd_mma.cu", "evd_cta.cu", "mle_kernels.cu", "post_kernels.cu"]
CUDA_LIB = os.path.join(LIBDIR, "libfringe_b200.so")
# profiling microbenchmarks: their own library (include/fringe_b200_prof.h), not part of the drop-in one
PROF_SOURCES = ["microbench.cu"]
PROF_LIB = os.path.join(LIBDIR, "libfringe_b200_prof.so")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd: list[str]) -> None:
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_cuda(force: bool = False, verbose_ptxas: bool = False, phase_clocks: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    psrcs = [os.path.join(CSRC, s) for s in PROF_SOURCES]
    deps = srcs + psrcs + [os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "fringe_b200.h"),
                           os.path.join(ROOT, "include", "fringe_b200_prof.h"), os.path.abspath(__file__)]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if force or _newer(CUDA_LIB, deps) or _newer(PROF_LIB, deps):
        flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
        if verbose_ptxas:
            flags += ["-Xptxas", "-v"]
        if phase_clocks:                     # profiling build: fringe_evd_phase_cycles
            flags += ["-DFRINGE_PHASE_CLOCKS"]
        flags += os.environ.get("FRINGE_NVCC_DEFS", "").split()      # A/B experiments only (extra -D switches)
        objdir = os.path.join(LIBDIR, "obj")
        os.makedirs(objdir, exist_ok=True)
        procs = []
        objs = []
        pobjs = []
        for s in srcs + psrcs:               # one nvcc per translation unit, in parallel
            o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
            (pobjs if s in psrcs else objs).append(o)
            cmd = [NVCC] + flags + ["-c", "-o", o, s]
            print("+", " ".join(cmd), flush=True)
            procs.append((cmd, subprocess.Popen(cmd)))
        for cmd, p in procs:
            if p.wait() != 0:
                raise subprocess.CalledProcessError(p.returncode, cmd)
        _run([NVCC] + ARCH + ["-shared", "-o", CUDA_LIB] + objs)
        _run([NVCC] + ARCH + ["-shared", "-o", PROF_LIB] + pobjs)
    return CUDA_LIB


def build_all(force: bool = False, phase_clocks: bool = False) -> None:
    build_cuda(force or phase_clocks, phase_clocks=phase_clocks)
    try:
        from . import build_host
    except ImportError:
        return
    build_host.build(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, phase_clocks="--phase-clocks" in sys.argv)
