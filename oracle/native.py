"""ctypes loader for ``oracle/csrc/oracle.c``.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

``ensure_built()`` compiles the library with the committed Makefile when it is missing
(``__graft_entry__.build()`` does so ahead of time, so the GPU box only loads it).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def ensure_built(force: bool = False) -> str:
    src = os.path.join(_HERE, "csrc", "oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(ensure_built())
        i64, f32p, i64p = ctypes.c_int64, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        _lib.oracle_fps_f32.argtypes = [f32p, i64, ctypes.c_int, i64, i64, i64p]
        _lib.oracle_ball_counts_f32.argtypes = [f32p, i64, ctypes.c_int, f32p, f32p, i64, i64p]
        _lib.oracle_min_dist_f32.argtypes = [f32p, i64, ctypes.c_int, f32p, i64, i64, f32p, f32p,
                                             ctypes.c_int, f32p]
        _lib.oracle_sample_points_f32.argtypes = [f32p, i64, ctypes.c_int, f32p, i64, ctypes.c_int, f32p]
        for fn in (_lib.oracle_fps_f32, _lib.oracle_ball_counts_f32, _lib.oracle_min_dist_f32,
                   _lib.oracle_sample_points_f32):
            fn.restype = ctypes.c_int
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def fps(points, n_samples: int, start_idx: int = 0) -> np.ndarray:
    p = _f32(points)
    n, d = p.shape
    n_samples = min(int(n_samples), n)
    out = np.empty(n_samples, dtype=np.int64)
    rc = lib().oracle_fps_f32(_fp(p), n, d, n_samples, int(start_idx),
                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    if rc != 0:
        raise RuntimeError(f"oracle_fps_f32 failed ({rc})")
    return out


def ball_counts(points, centers, radii) -> np.ndarray:
    p, c, r = _f32(points), _f32(centers), _f32(radii)
    out = np.empty(c.shape[0], dtype=np.int64)
    lib().oracle_ball_counts_f32(_fp(p), p.shape[0], p.shape[1], _fp(c), _fp(r), c.shape[0],
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return out


def min_dist(points, samples, centers=None, radii=None) -> np.ndarray:
    """(S,R) float32 brute-force min distance, restricted to the balls when given."""
    p, x = _f32(points), _f32(samples)
    S, R, d = x.shape
    use_ball = centers is not None
    c = _f32(centers) if use_ball else np.zeros((S, d), np.float32)
    r = _f32(radii) if use_ball else np.zeros((S,), np.float32)
    out = np.empty((S, R), dtype=np.float32)
    rc = lib().oracle_min_dist_f32(_fp(p), p.shape[0], d, _fp(x), S, R, _fp(c), _fp(r),
                                   int(use_ball), _fp(out))
    if rc != 0:
        raise RuntimeError(f"oracle_min_dist_f32 failed ({rc})")
    return out


def sample_points(weights, vertices) -> np.ndarray:
    w, v = _f32(weights), _f32(vertices)
    S, K, d = v.shape
    out = np.empty((S, w.shape[0], d), dtype=np.float32)
    lib().oracle_sample_points_f32(_fp(w), w.shape[0], K, _fp(v), S, d, _fp(out))
    return out
