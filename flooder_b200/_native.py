"""Loader for the native pieces.  There is NO fallback: if the CUDA library or the torch
extension is missing, importing the product path raises."""
from __future__ import annotations

import ctypes
import importlib.machinery
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(BUILD, "libflood_b200.so")
EXT_PATH = os.path.join(BUILD, "_flood_ext.so")

_ext = None
_lib = None


class NativeLibraryMissing(ImportError):
    pass


def _require(path: str) -> str:
    if not os.path.exists(path):
        raise NativeLibraryMissing(
            f"{path} is missing: build it with `python -m flooder_b200.build` "
            "(nvcc, sm_100a).  flooder_b200 has no CPU or pure-PyTorch fallback."
        )
    return path


def ext():
    """The PyTorch C++ extension (product path)."""
    global _ext
    if _ext is None:
        import torch  # noqa: F401  (libtorch must be loaded first)

        _require(LIB_PATH)
        path = _require(EXT_PATH)
        loader = importlib.machinery.ExtensionFileLoader("_flood_ext", path)
        spec = importlib.util.spec_from_loader("_flood_ext", loader, origin=path)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        if mod.abi_version() != 1:
            raise ImportError(f"libflood_b200 ABI {mod.abi_version()} != 1")
        # FLOODER_B200_OPTIONS="name=value,..." presets library options (experiments / diagnostics)
        for item in filter(None, os.environ.get("FLOODER_B200_OPTIONS", "").split(",")):
            name, _, value = item.partition("=")
            mod.set_option(name.strip(), int(value))
        _ext = mod
    return _ext


def cdll() -> ctypes.CDLL:
    """The bare C ABI through ctypes (used by the ABI tests and by INTEGRATION.md's example)."""
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_require(LIB_PATH))
        c_i64, c_int, c_vp, c_sz = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
        lib.flood_abi_version.restype = c_int
        lib.flood_last_error.restype = ctypes.c_char_p
        lib.flood_device_info.argtypes = [ctypes.POINTER(c_int), ctypes.POINTER(c_int)]
        lib.flood_fps_workspace_bytes.argtypes = [c_i64, c_int, c_i64]
        lib.flood_fps_workspace_bytes.restype = c_sz
        lib.flood_fps_f32.argtypes = [c_vp, c_i64, c_int, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]
        lib.flood_fps_grid_f32.argtypes = [c_vp, c_vp, c_i64, c_int, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]
        lib.flood_cloud_workspace_bytes.argtypes = [c_i64, c_int]
        lib.flood_cloud_workspace_bytes.restype = c_sz
        lib.flood_cloud_build_f32.argtypes = [c_vp, c_i64, c_int, c_int, c_vp, c_sz, c_vp]
        lib.flood_bounding_balls_f32.argtypes = [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp]
        lib.flood_covering_workspace_bytes.argtypes = [c_i64, c_i64, c_int]
        lib.flood_covering_workspace_bytes.restype = c_sz
        lib.flood_covering_radius_f32.argtypes = [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, c_i64,
                                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]
        lib.flood_covering_bricks.argtypes = [c_i64, c_int, c_vp, c_int, ctypes.POINTER(c_int)]
        lib.flood_covering_plan_f32.argtypes = [c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_vp, c_vp]
        lib.flood_face_max_f32.argtypes = [c_vp, c_i64, c_i64, c_vp, c_int, c_vp, c_vp]
        lib.flood_set_option.argtypes = [ctypes.c_char_p, c_int]
        lib.flood_launch_count.argtypes = [c_int]
        lib.flood_launch_count.restype = ctypes.c_longlong
        for fn in ("flood_device_info", "flood_fps_f32", "flood_fps_grid_f32", "flood_cloud_build_f32", "flood_bounding_balls_f32",
                   "flood_covering_radius_f32", "flood_covering_plan_f32", "flood_covering_bricks", "flood_face_max_f32",
                   "flood_set_option"):
            getattr(lib, fn).restype = c_int
        _lib = lib
    return _lib


EXPORTED_SYMBOLS = [
    "flood_abi_version", "flood_last_error", "flood_device_info", "flood_set_option", "flood_kernel_ms",
    "flood_launch_count",
    "flood_fps_workspace_bytes", "flood_fps_f32", "flood_fps_grid_f32",
    "flood_cloud_workspace_bytes", "flood_cloud_build_f32",
    "flood_bounding_balls_f32",
    "flood_covering_workspace_bytes", "flood_covering_radius_f32", "flood_covering_plan_f32",
    "flood_covering_bricks",
    "flood_bounding_balls_f64", "flood_covering_workspace_bytes_f64", "flood_covering_radius_f64", "flood_face_max_f64",
    "flood_face_max_f32",
]
