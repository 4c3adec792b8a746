"""World-size-2 CPU (gloo) test of the data-parallel decomposition the multi-GPU path relies on:
each rank owns B/G rows and chains of every mini-batch (bench.shard_rows), statistics are
summed with one all-reduce, every rank applies the same update.  Computed here with the
oracle's arithmetic on the CPU: the sharded + all-reduced run must equal the single-process
run, including the chain states (random numbers are keyed by the GLOBAL row)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

V, H, B, K, NB = 12, 8, 8, 2, 3


def _pcd_batched(W, a, b, X, cv, ch, U_h, U_v, lr, B_global, reduce):
    """One PCD step in the batched form on a row shard; `reduce` sums the statistics over ranks."""
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    for s in range(K):
        ch = (U_h[s] < sig(b + cv @ W)).astype(np.float64)
        cv = (U_v[s] < sig(a + ch @ W.T)).astype(np.float64)
    hd = sig(b + X @ W)
    flat = np.concatenate([(X.T @ hd - cv.T @ ch).ravel(), (X - cv).sum(0), (hd - ch).sum(0)])
    flat = reduce(flat)
    dW = flat[:V * H].reshape(V, H)
    W = W + dW * (lr / B_global)
    a = a + flat[V * H:V * H + V] * (lr / B_global)
    b = b + flat[V * H + V:] * (lr / B_global)
    return W, a, b, cv, ch


def _problem():
    from tests.philox_ref import uniforms
    rng = np.random.default_rng(3)
    X = rng.integers(0, 2, (NB * B, V)).astype(np.float64)
    W = 0.3 * rng.standard_normal((V, H))
    seed = 42
    cv, ch = uniforms(seed, 0, B, V), uniforms(seed, 1, B, H)
    U_h = [np.stack([uniforms(seed, 2 + 2 * (t * K + s), B, H) for s in range(K)]) for t in range(NB)]
    U_v = [np.stack([uniforms(seed, 3 + 2 * (t * K + s), B, V) for s in range(K)]) for t in range(NB)]
    return X, W, cv, ch, U_h, U_v


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bench import shard_rows
    X, W, cv, ch, U_h, U_v = _problem()
    Bl = B // world
    sl = slice(rank * Bl, (rank + 1) * Bl)
    Xl = shard_rows(X, B, NB, rank, world)
    a, b = np.zeros(V), np.zeros(H)
    cv, ch = cv[sl], ch[sl]

    def reduce(flat):
        t = torch.from_numpy(flat.copy())
        dist.all_reduce(t)
        return t.numpy()

    for t in range(NB):
        W, a, b, cv, ch = _pcd_batched(W, a, b, Xl[t * Bl:(t + 1) * Bl], cv, ch, U_h[t][:, sl], U_v[t][:, sl], 0.1, B, reduce)
    out = [None] * world
    dist.all_gather_object(out, (W, a, b, cv, ch))
    if rank == 0:
        ret.put(out)
    dist.destroy_process_group()


def test_sharded_allreduce_equals_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29611, ret)) for r in range(world)]
    for p in procs:
        p.start()
    out = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X, W, cv, ch, U_h, U_v = _problem()
    a, b = np.zeros(V), np.zeros(H)
    for t in range(NB):
        W, a, b, cv, ch = _pcd_batched(W, a, b, X[t * B:(t + 1) * B], cv, ch, U_h[t], U_v[t], 0.1, B, lambda f: f)
    for r in range(world):
        assert np.array_equal(out[r][0], out[0][0])  # replicas stay bit-identical
    assert np.allclose(out[0][0], W, rtol=0, atol=1e-12) and np.allclose(out[0][1], a, atol=1e-12)
    assert np.array_equal(np.concatenate([o[3] for o in out]), cv)
    assert np.array_equal(np.concatenate([o[4] for o in out]), ch)


def test_shard_rows_layout():
    from bench import shard_rows
    A = np.arange(2 * 8 * 3).reshape(16, 3)
    s0, s1 = shard_rows(A, 8, 2, 0, 2), shard_rows(A, 8, 2, 1, 2)
    assert np.array_equal(s0, np.concatenate([A[0:4], A[8:12]])) and np.array_equal(s1, np.concatenate([A[4:8], A[12:16]]))
