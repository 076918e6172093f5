"""Builds libqcknot.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every translation unit is compiled to an object under build/ (only when it or a header changed, in parallel) and the
objects are linked into quantumcollocation.jl_b200/libqcknot.so."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["csrc/qck_kernels.cu", "csrc/qck_rowslice.cu", "csrc/qck_rs3.cu", "csrc/qck_column.cu", "csrc/qck_pack.cu",
           "csrc/qck_host.cpp", "csrc/qck_pipe.cpp", "csrc/qck_multi.cpp"]
HEADERS = ["csrc/qck_internal.h", "csrc/qck_handle.h", "csrc/qck_device.cuh", "../include/qcknot.h"]
LIB = os.path.join(HERE, "libqcknot.so")
OBJDIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-diag-suppress", "177",
]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]


def _obj(src: str) -> str:
    return os.path.join(OBJDIR, os.path.basename(src) + ".o")


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
    srcs = _sources()
    todo = [s for s in srcs if force or _stale(_obj(s), [s] + HEADERS)]

    def compile_one(src):
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", _obj(src), src]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        return src, subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        for src, res in ex.map(compile_one, todo):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
            if verbose:
                print(f"== {src}\n" + res.stdout + res.stderr)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB,
            *[_obj(s) for s in srcs], "-ldl", "-lpthread"]
    res = subprocess.run(link, cwd=HERE, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
