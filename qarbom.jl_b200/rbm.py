"""Host-side model types with the reference's names and fields (src/rbm.jl:2-39, 43-99)."""
from __future__ import annotations

from typing import Optional

import numpy as np

_default_rng = np.random.default_rng()


class AbstractRBM:
    gaussian = False
    n_classifiers = 0

    def _init(self, n_visible, n_hidden, W, rng):
        rng = rng or _default_rng
        self.n_visible, self.n_hidden = int(n_visible), int(n_hidden)
        # `W = randn(n_visible, n_hidden)` (sigma = 1, rbm.jl:44) or `copy(W)` (rbm.jl:53)
        self.W = rng.standard_normal((n_visible, n_hidden)) if W is None else np.array(W, dtype=np.float64).reshape(
            n_visible, n_hidden).copy()
        self.a = np.zeros(n_visible)
        self.b = np.zeros(n_hidden)


class RBM(AbstractRBM):
    def __init__(self, n_visible: int, n_hidden: int, W: Optional[np.ndarray] = None, rng=None):
        self._init(n_visible, n_hidden, W, rng)


class GRBM(AbstractRBM):
    gaussian = True

    def __init__(self, n_visible: int, n_hidden: int, W: Optional[np.ndarray] = None, rng=None):
        self._init(n_visible, n_hidden, W, rng)


class RBMClassifier(AbstractRBM):
    def __init__(self, n_visible: int, n_hidden: int, n_classifiers: int, W=None, U=None, rng=None):
        rng = rng or _default_rng
        self._init(n_visible, n_hidden, W, rng)
        self.n_classifiers = int(n_classifiers)
        self.U = rng.standard_normal((n_classifiers, n_hidden)) if U is None else np.array(U, dtype=np.float64).reshape(
            n_classifiers, n_hidden).copy()
        self.c = np.zeros(n_classifiers)
        if self.gaussian and W is not None:  # GRBMClassifier 5-arg ctor, rbm.jl:94-99
            self.a[:] = 1e-5
            self.b[:] = 1e-5
            self.c[:] = 1e-5


class GRBMClassifier(RBMClassifier):
    gaussian = True


def copy_rbm(rbm: AbstractRBM) -> dict:
    """Snapshot of every parameter array (copy_rbm, rbm.jl:231-233,244-246)."""
    out = {"W": rbm.W.copy(), "a": rbm.a.copy(), "b": rbm.b.copy()}
    if rbm.n_classifiers:
        out["U"], out["c"] = rbm.U.copy(), rbm.c.copy()
    return out


def copy_rbm_into(snap: dict, rbm: AbstractRBM) -> None:
    """copy_rbm!(src, target), rbm.jl:235-255."""
    for k, v in snap.items():
        getattr(rbm, k)[...] = v
