"""ORACLE (test infrastructure, NOT product code): ctypes front-end of oracle/tokenizer_ref.c plus
the numpy FPS restatement the reference's own test uses
(third_party/torkit3d/tests/ops/test_sample_farthest_points.py:7-38)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "tokenizer_ref.c")
_LIB = os.path.join(_HERE, "_build", "liboracle_tokenizer.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _LIB, _SRC, "-lm"])
    return _LIB


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i64, p = ctypes.c_int64, ctypes.c_void_p
        for name in ("oracle_fps_f32", "oracle_fps_closed_f32"):
            fn = getattr(_lib, name)
            fn.argtypes = [p, i64, i64, i64, p]
            fn.restype = ctypes.c_int
        _lib.oracle_knn_f32.argtypes = [p, p, i64, i64, i64, i64, p, p]
        _lib.oracle_knn_f32.restype = ctypes.c_int
    return _lib


def _fps(fn_name: str, points: np.ndarray, num_samples: int) -> np.ndarray:
    pts = np.ascontiguousarray(points, dtype=np.float32)
    assert pts.ndim == 3 and pts.shape[2] == 3
    B, N, _ = pts.shape
    out = np.empty((B, num_samples), dtype=np.int64)
    rc = getattr(_load(), fn_name)(pts.ctypes.data, B, N, num_samples, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle fps: invalid arguments")
    return out


def fps(points: np.ndarray, num_samples: int) -> np.ndarray:
    """Literal simulation of the reference CUDA kernel (strided scan + shared-memory tree)."""
    return _fps("oracle_fps_f32", points, num_samples)


def fps_closed(points: np.ndarray, num_samples: int) -> np.ndarray:
    """Closed form of the same tie-break rule."""
    return _fps("oracle_fps_closed_f32", points, num_samples)


def fps_numpy(points: np.ndarray, num_samples: int) -> np.ndarray:
    """The reference test's numpy oracle (np.argmax = first maximum; dtype of the input)."""
    index = []
    for pts in points:
        idx, cur, d2s = [0], 0, None
        for _ in range(1, num_samples):
            d = np.square(pts - pts[cur][None, :]).sum(1)
            d2s = d if d2s is None else np.minimum(d, d2s)
            cur = int(np.argmax(d2s))
            idx.append(cur)
        index.append(idx)
    return np.asarray(index)


def knn(query: np.ndarray, key: np.ndarray, k: int):
    """Exact K nearest keys per query, sorted by (squared distance, index)."""
    q = np.ascontiguousarray(query, dtype=np.float32)
    kk = np.ascontiguousarray(key, dtype=np.float32)
    B, Q, _ = q.shape
    N = kk.shape[1]
    idx = np.empty((B, Q, k), dtype=np.int64)
    d2 = np.empty((B, Q, k), dtype=np.float32)
    rc = _load().oracle_knn_f32(q.ctypes.data, kk.ctypes.data, B, Q, N, k, idx.ctypes.data, d2.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle knn: invalid arguments")
    return idx, d2
