"""ctypes wrapper around the C restatement (oracle/qarbom_oracle.c).

TEST INFRASTRUCTURE ONLY (see the header of qarbom_oracle.py).  `build()` compiles the C
file with gcc; nothing here touches /root/reference at run time.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "qarbom_oracle.c")
_lib = None

_D = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.qo_pcd_epoch.restype = ctypes.c_int
        _lib.qo_pcd_epoch.argtypes = [_i64, _i64, _i64, ctypes.c_int, _D, _D, _D, _D, _D, _D, _D, _i64, _i64, _i64, _i64,
                                      ctypes.c_double, ctypes.c_double, _D, _D, _D, _D, _i64, _D, _i64,
                                      ctypes.c_uint64, _i64, _i64, ctypes.c_int]
        _lib.qo_cd_epoch.restype = ctypes.c_int
        _lib.qo_cd_epoch.argtypes = [_i64, _i64, _i64, ctypes.c_int, _D, _D, _D, _D, _D, _D, _D, _i64, _i64, _i64,
                                     ctypes.c_double, ctypes.c_double, _D, _i64, _D, _i64, ctypes.c_uint64,
                                     _i64, _i64, ctypes.c_int]
        _lib.qo_fast_pcd_epoch.restype = ctypes.c_int
        _lib.qo_fast_pcd_epoch.argtypes = [_i64, _i64, _i64, ctypes.c_int, _D, _D, _D, _D, _D, _D, _D, _i64, _i64, _i64,
                                           ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _D, _D, _D,
                                           _D, _i64, _D, _i64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
        _lib.qo_eval_mse.restype = ctypes.c_double
        _lib.qo_eval_mse.argtypes = [_i64, _i64, ctypes.c_int, _D, _D, _D, _D, _i64, _D, _i64, ctypes.c_uint64,
                                     ctypes.c_int]
        _lib.qo_mean_free_energy.restype = ctypes.c_double
        _lib.qo_mean_free_energy.argtypes = [_i64, _i64, _i64, ctypes.c_int, _D, _D, _D, _D, _D, _D, _D, _i64]
        _lib.qo_max_threads.restype = ctypes.c_int
    return _lib


def _p(a: Optional[np.ndarray]):
    if a is None:
        return ctypes.cast(None, _D)
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_D)


class CModel:
    """Parameters in the C oracle's layout (column-major W: stored as W.T C-contiguous)."""

    def __init__(self, W: np.ndarray, a, b, U=None, c=None, gaussian=False):
        self.V, self.H = W.shape
        self.L = 0 if U is None else U.shape[0]
        self.Wc = np.ascontiguousarray(np.asarray(W, dtype=np.float64).T)        # [H][V] == col-major V x H
        self.a = np.array(a, dtype=np.float64)
        self.b = np.array(b, dtype=np.float64)
        self.Uc = None if U is None else np.ascontiguousarray(np.asarray(U, dtype=np.float64).T)
        self.c = None if c is None else np.array(c, dtype=np.float64)
        self.gaussian = int(bool(gaussian))

    @property
    def W(self):
        return self.Wc.T

    @property
    def U(self):
        return None if self.Uc is None else self.Uc.T


def pcd_epoch(m: CModel, X, Y, B, k, lr, llr, chv, chh, chy, u=None, z=None, seed=0, first_batch=0,
              n_batches=-1, threads=1, n_chains=None) -> int:
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = None if Y is None else np.ascontiguousarray(Y, dtype=np.float64)
    u = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
    z = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
    return lib().qo_pcd_epoch(m.V, m.H, m.L, m.gaussian, _p(m.Wc), _p(m.a), _p(m.b), _p(m.Uc), _p(m.c), _p(X), _p(Y),
                              X.shape[0], B, B if n_chains is None else n_chains, k, lr, llr, _p(chv), _p(chh), _p(chy),
                              _p(u), 0 if u is None else u.size, _p(z), 0 if z is None else z.size,
                              seed, first_batch, n_batches, threads)


def fast_pcd_epoch(m: CModel, X, Y, B, k, lr, flr, llr, fllr, chv, chh, chy, u=None, z=None, seed=0, pair_chains=None,
                   threads=1) -> int:
    """One epoch of fast_persistent_contrastive_divergence! (src/training/train_fast_pcd.jl:1-127)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = None if Y is None else np.ascontiguousarray(Y, dtype=np.float64)
    u = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
    z = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
    return lib().qo_fast_pcd_epoch(m.V, m.H, m.L, m.gaussian, _p(m.Wc), _p(m.a), _p(m.b), _p(m.Uc), _p(m.c), _p(X), _p(Y),
                                   X.shape[0], B, k, lr, flr, llr, fllr, _p(chv), _p(chh), _p(chy),
                                   _p(u), 0 if u is None else u.size, _p(z), 0 if z is None else z.size, seed,
                                   -1 if pair_chains is None else int(bool(pair_chains)), threads)


def cd_epoch(m: CModel, X, Y, B, k, lr, llr, u=None, z=None, seed=0, first_batch=0, n_batches=-1, threads=1) -> int:
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = None if Y is None else np.ascontiguousarray(Y, dtype=np.float64)
    u = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
    z = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
    return lib().qo_cd_epoch(m.V, m.H, m.L, m.gaussian, _p(m.Wc), _p(m.a), _p(m.b), _p(m.Uc), _p(m.c), _p(X), _p(Y),
                             X.shape[0], B, k, lr, llr, _p(u), 0 if u is None else u.size,
                             _p(z), 0 if z is None else z.size, seed, first_batch, n_batches, threads)


def eval_mse(m: CModel, X, z=None, seed=0, threads=1) -> float:
    X = np.ascontiguousarray(X, dtype=np.float64)
    z = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
    return lib().qo_eval_mse(m.V, m.H, m.gaussian, _p(m.Wc), _p(m.a), _p(m.b), _p(X), X.shape[0],
                             _p(z), 0 if z is None else z.size, seed, threads)


def mean_free_energy(m: CModel, X, Y=None) -> float:
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = None if Y is None else np.ascontiguousarray(Y, dtype=np.float64)
    return lib().qo_mean_free_energy(m.V, m.H, m.L, m.gaussian, _p(m.Wc), _p(m.a), _p(m.b), _p(m.Uc), _p(m.c),
                                     _p(X), _p(Y), X.shape[0])


def max_threads() -> int:
    return lib().qo_max_threads()
