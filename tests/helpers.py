"""Shared helpers for the parity tests (oracle side bookkeeping, bf16 rounding, streams)."""
from __future__ import annotations

import numpy as np

from oracle import qarbom_oracle as qo


def u24(rng, *shape):
    """Uniforms exactly representable in fp32 (k / 2^24)."""
    return rng.integers(0, 1 << 24, size=shape).astype(np.float64) / float(1 << 24)


def bf16_round(x):
    """Round-to-nearest-even to bfloat16, returned as float64."""
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) & np.uint64(0xFFFF0000)
    return u.astype(np.uint32).view(np.float32).astype(np.float64).reshape(np.shape(x))


def f32_round(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def make_oracle_model(rng, V, H, L=0, gaussian=False, scale=0.3, bias_scale=0.1, bf16=False):
    W = f32_round(rng.standard_normal((V, H)) * scale)
    U = f32_round(rng.standard_normal((L, H)) * scale) if L else None
    if bf16:
        W = bf16_round(W)
        U = bf16_round(U) if L else None
    m = qo.make_model(V, H, L, gaussian, W=W, U=U)
    m.a[:] = f32_round(rng.standard_normal(V) * bias_scale)
    m.b[:] = f32_round(rng.standard_normal(H) * bias_scale)
    if L:
        m.c[:] = f32_round(rng.standard_normal(L) * bias_scale)
    return m


def to_host_rbm(q, m):
    """oracle Model -> qarbom_b200 host model of the matching type."""
    L = m.n_classifiers
    if L:
        cls = q.GRBMClassifier if m.gaussian else q.RBMClassifier
        r = cls(m.n_visible, m.n_hidden, L, W=m.W, U=m.U)
        r.c[:] = m.c
    else:
        cls = q.GRBM if m.gaussian else q.RBM
        r = cls(m.n_visible, m.n_hidden, W=m.W)
    r.a[:] = m.a
    r.b[:] = m.b
    return r


class DenseStreams:
    """Dense per-step injection arrays (the C-ABI layout) <-> the reference's serial order."""

    def __init__(self, rng, k, B, V, H, L=0, gaussian=False):
        self.k, self.B, self.V, self.H, self.L, self.gaussian = k, B, V, H, L, gaussian
        self.U_h = u24(rng, k, B, H)
        self.U_v = None if gaussian else u24(rng, k, B, V)
        self.U_y = u24(rng, k, B, L) if L else None
        self.Z_v = f32_round(rng.standard_normal((k, B, V))) if gaussian else None

    def _order(self, order):
        k, B = self.k, self.B
        if order == "pcd":   # step-major, chain-minor (fantasy_data.jl:12-29)
            return [(s, i) for s in range(k) for i in range(B)]
        return [(s, i) for i in range(B) for s in range(k)]  # cd: sample-major (train_cd.jl:4-19)

    def serial(self, order):
        parts = []
        for s, i in self._order(order):
            parts.append(self.U_h[s, i])
            if self.U_v is not None:
                parts.append(self.U_v[s, i])
            if self.U_y is not None:
                parts.append(self.U_y[s, i])
        u = np.concatenate(parts)
        z = None
        if self.gaussian:
            z = np.concatenate([self.Z_v[s, i] for s, i in self._order(order)])
        return u, z

    def absorb(self, serial_u, order):
        """Write a (guard-nudged) serial uniform stream back into the dense arrays."""
        pos = 0
        for s, i in self._order(order):
            self.U_h[s, i] = serial_u[pos:pos + self.H]; pos += self.H
            if self.U_v is not None:
                self.U_v[s, i] = serial_u[pos:pos + self.V]; pos += self.V
            if self.U_y is not None:
                self.U_y[s, i] = serial_u[pos:pos + self.L]; pos += self.L
        assert pos == serial_u.size


def chains_from_arrays(v, h, y=None):
    return [qo.FantasyData(v[i].copy(), h[i].copy(), None if y is None else y[i].copy()) for i in range(len(v))]


def chains_to_arrays(chains):
    v = np.stack([c.v for c in chains]); h = np.stack([c.h for c in chains])
    y = None if chains[0].y is None else np.stack([c.y for c in chains])
    return v, h, y


def rel_err(a, b, atol=2e-7):
    """||a - b|| / ||b||, with an fp32-rounding floor (atol per element) so that exactly-zero
    increments (e.g. v_model == v_data on a 3x2 model) compare as equal."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = float(np.linalg.norm(a - b))
    if d <= atol * np.sqrt(a.size):
        return 0.0
    return d / max(float(np.linalg.norm(b)), 1e-300)
