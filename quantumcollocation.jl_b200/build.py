"""Builds libqcknot.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every translation unit is compiled to an object under build/ (only when it or a header changed, in parallel) and the
objects are linked into quantumcollocation.jl_b200/libqcknot.so."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
# (source, extra flags, object tag): the rs3 kernel file is compiled once per drive count, in parallel
UNITS = [("csrc/qck_kernels.cu", [], ""), ("csrc/qck_rowslice.cu", [], ""), ("csrc/qck_column.cu", [], ""), ("csrc/qck_big.cu", [], ""), ("csrc/qck_expeig.cu", [], ""), ("csrc/qck_colexp.cu", [], ""), ("csrc/qck_genexp.cu", [], ""), ("csrc/qck_pack.cu", [], ""),
         ("csrc/qck_objective.cu", [], ""), ("csrc/qck_rollout.cu", [], ""),
         ("csrc/qck_rs3.cu", ["-DQCK_RS3_ND=1"], ".nd1"), ("csrc/qck_rs3.cu", ["-DQCK_RS3_ND=2"], ".nd2"),
         ("csrc/qck_rs3.cu", ["-DQCK_RS3_ND=3"], ".nd3"), ("csrc/qck_rs3.cu", ["-DQCK_RS3_ND=4"], ".nd4"),
         ("csrc/qck_host.cpp", [], ""), ("csrc/qck_pipe.cpp", [], ""), ("csrc/qck_multi.cpp", [], "")]
SOURCES = sorted({u[0] for u in UNITS})
HEADERS = ["csrc/qck_internal.h", "csrc/qck_handle.h", "csrc/qck_device.cuh", "../include/qcknot.h"]
LIB = os.path.join(HERE, "libqcknot.so")
OBJDIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-diag-suppress", "177",
]


def _sources():
    return SOURCES


def _obj(unit) -> str:
    return os.path.join(OBJDIR, os.path.basename(unit[0]) + unit[2] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(HERE, d)) > t for d in deps)


def needs_build() -> bool:
    return _stale(LIB, _sources() + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJDIR, exist_ok=True)
    todo = [u for u in UNITS if force or _stale(_obj(u), [u[0]] + HEADERS)]

    def compile_one(unit):
        cmd = [nvcc, *NVCC_FLAGS, *unit[1], "-c", "-o", _obj(unit), unit[0]]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        return unit[0] + unit[2], subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        for src, res in ex.map(compile_one, todo):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
            if verbose:
                print(f"== {src}\n" + res.stdout + res.stderr)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB,
            *[_obj(u) for u in UNITS], "-ldl", "-lpthread"]
    res = subprocess.run(link, cwd=HERE, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
