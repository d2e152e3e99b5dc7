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

CU_SOURCES = ["abi.cu", "cloud.cu", "balls.cu", "covering.cu", "fps.cu", "f64.cu"] + [f"covering_d{d}.cu" for d in range(1, 9)]
CU_HEADERS = ["common.cuh", "covering_kernels.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJ = os.path.join(OUT, "obj")


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
    """One object per .cu (compiled in parallel: the evaluation kernel is instantiated per ambient
    dimension in its own translation unit), then one link step."""
    from concurrent.futures import ThreadPoolExecutor

    headers = [os.path.join(CSRC, h) for h in CU_HEADERS] + [os.path.join(INCLUDE, "flood_b200.h")]
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for name in CU_SOURCES:
        src = os.path.join(CSRC, name)
        obj = os.path.join(OBJ, name[:-3] + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    if jobs:
        def run(cmd):
            res = subprocess.run(cmd, capture_output=True, text=True)
            return cmd, res
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            for cmd, res in pool.map(run, jobs):
                if verbose or res.returncode != 0:
                    print(" ".join(cmd))
                    print(res.stdout + res.stderr)
                if res.returncode != 0:
                    raise subprocess.CalledProcessError(res.returncode, cmd)
    objs = [os.path.join(OBJ, name[:-3] + ".o") for name in CU_SOURCES]
    if force or jobs or _stale(LIB, objs):
        subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs], check=True)
    return LIB


def _ext_command(cxx: str, src: str):
    import torch
    from torch.utils import cpp_extension

    incs = cpp_extension.include_paths("cuda") + [sysconfig.get_paths()["include"], INCLUDE]
    libdirs = cpp_extension.library_paths("cuda")
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
    return cmd


_SELF_TEST = """
import importlib.machinery, importlib.util, sys
import torch
loader = importlib.machinery.ExtensionFileLoader("_flood_ext", sys.argv[1])
mod = importlib.util.module_from_spec(importlib.util.spec_from_loader("_flood_ext", loader, origin=sys.argv[1]))
loader.exec_module(mod)
try:
    mod._raise_for_code(-4)
except RuntimeError as exc:
    assert "self-test failed (-4)" in str(exc)
    print("ok")
"""


def _ext_self_test() -> bool:
    """A library error code must surface as a Python RuntimeError.  (A toolchain whose start files /
    libgcc differ from the system's has produced an extension that crashed on that path: the build
    tries the next compiler instead of shipping it.)"""
    res = subprocess.run([sys.executable, "-c", _SELF_TEST, EXT], capture_output=True, text=True)
    return res.returncode == 0 and "ok" in res.stdout


def build_ext(force: bool = False, verbose: bool = False) -> str:
    src = os.path.join(CSRC, "torch_binding.cpp")
    if not (force or _stale(EXT, [src, os.path.join(INCLUDE, "flood_b200.h"), LIB])):
        return EXT
    os.makedirs(OUT, exist_ok=True)
    candidates = []
    for cxx in (os.environ.get("FLOOD_CXX"), "/usr/bin/g++", shutil.which("g++"), os.environ.get("CXX")):
        if cxx and os.path.exists(cxx) and cxx not in candidates:
            candidates.append(cxx)
    for cxx in candidates:
        cmd = _ext_command(cxx, src)
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        if _ext_self_test():
            return EXT
        print(f"flooder_b200.build: extension built with {cxx} fails its error-path self-test, trying the next compiler",
              file=sys.stderr)
    raise RuntimeError("no host compiler produced a working _flood_ext.so")


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_lib(force, verbose)
    build_ext(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv)
    print("built:", LIB, EXT)
