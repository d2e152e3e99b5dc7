"""In-tree build of the native pieces (sm_100a only).

    python -m flooder_b200.build            # build what is stale
    python -m flooder_b200.build --force

Outputs (git-ignored, shipped to the GPU box by gpurun):
    flooder_b200/_build/libflood_b200.so    CUDA kernels + C ABI (include/flood_b200.h), nvcc
    flooder_b200/_build/_flood_ext.so       PyTorch C++ extension that forwards tensors to the C ABI
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
OUT = os.path.join(HERE, "_build")
INCLUDE = os.path.join(ROOT, "include")

LIB = os.path.join(OUT, "libflood_b200.so")
EXT = os.path.join(OUT, "_flood_ext.so")

CU_SOURCES = ["abi.cu", "cloud.cu", "balls.cu", "covering.cu", "fps.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--threads", "4", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(INCLUDE, "flood_b200.h")]
    if force or _stale(LIB, deps):
        os.makedirs(OUT, exist_ok=True)
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-o", LIB, *srcs]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


def build_ext(force: bool = False, verbose: bool = False) -> str:
    src = os.path.join(CSRC, "torch_binding.cpp")
    if not (force or _stale(EXT, [src, os.path.join(INCLUDE, "flood_b200.h"), LIB])):
        return EXT
    import torch
    from torch.utils import cpp_extension

    os.makedirs(OUT, exist_ok=True)
    incs = cpp_extension.include_paths("cuda") + [sysconfig.get_paths()["include"], INCLUDE]
    libdirs = cpp_extension.library_paths("cuda")
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_flood_ext",
           "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    for i in incs:
        cmd += ["-isystem", i]
    cmd += [src, "-o", EXT]
    for d in libdirs:
        cmd += [f"-L{d}"]
    cmd += [f"-L{OUT}", "-lflood_b200", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
            "-ltorch_python", "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return EXT


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_lib(force, verbose)
    build_ext(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv)
    print("built:", LIB, EXT)
