"""World-size-2 CPU (gloo) tests of the data-parallel decompositions the multi-GPU path relies on:
each rank owns B/G rows and chains of every mini-batch (bench.shard_rows); the statistics are either
summed with one all-reduce and every rank applies the same update, or reduce-scattered over the hidden
units with a shard update and an all-gather of the weights ("sharded_update"), or exchanged as bit-packed
chain states ("state_exchange").  Computed here with the
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


def _allreduce(x):
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
    dist.all_reduce(t)
    return t.numpy()


def _allgather(x, world):
    t = torch.from_numpy(np.ascontiguousarray(x).copy())
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.numpy() for o in out]


def _pcd_sharded_update(W, a, b, X, cv, ch, U_h, U_v, lr, B_global, rank, world, state_exchange):
    """The two exchanges of the bf16 multi-GPU path, with the oracle's arithmetic: every rank owns H / world hidden
    units.  state_exchange = False: reduce-scatter of the gradient, update of the owned columns, all-gather of the
    weights ("sharded_update").  True: only the POSITIVE product is reduce-scattered (the device does that behind the
    chain advance); the binary chain states are all-gathered as bits and every rank computes its own shard of the full
    negative product ("state_exchange")."""
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    hs = slice(rank * (H // world), (rank + 1) * (H // world))
    hd = sig(b + X @ W)                                   # positive phase first: every GEMM of the step uses W_t
    for s in range(K):
        ch = (U_h[s] < sig(b + cv @ W)).astype(np.float64)
        cv = (U_v[s] < sig(a + ch @ W.T)).astype(np.float64)
    if state_exchange:
        pos = _allreduce(X.T @ hd)[:, hs]                 # reduce-scatter: keep the owned hidden columns
        bits = _allgather(np.packbits(np.concatenate([cv, ch], axis=1).astype(np.uint8), axis=0), world)   # bits along the rows
        both = np.concatenate([np.unpackbits(x, axis=0)[:cv.shape[0]] for x in bits]).astype(np.float64)
        cv_all, ch_all = both[:, :V], both[:, V:]
        g = pos - cv_all.T @ ch_all[:, hs]
    else:
        g = _allreduce(X.T @ hd - cv.T @ ch)[:, hs]
    bias = _allreduce(np.concatenate([(X - cv).sum(0), (hd - ch).sum(0)]))
    W = np.concatenate(_allgather(W[:, hs] + g * (lr / B_global), world), axis=1)   # shard update, then all-gather of the weights
    a = a + bias[:V] * (lr / B_global)
    b = b + bias[V:] * (lr / B_global)
    return W, a, b, cv, ch


def _worker(rank, world, port, ret, mode="allreduce"):
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
        if mode == "allreduce":
            W, a, b, cv, ch = _pcd_batched(W, a, b, Xl[t * Bl:(t + 1) * Bl], cv, ch, U_h[t][:, sl], U_v[t][:, sl], 0.1, B, reduce)
        else:
            # (the uniform-float chains of _init_fantasy_data become binary in the first advance, before they are packed)
            W, a, b, cv, ch = _pcd_sharded_update(W, a, b, Xl[t * Bl:(t + 1) * Bl], cv, ch, U_h[t][:, sl], U_v[t][:, sl], 0.1, B,
                                                  rank, world, mode == "state_exchange")
    out = [None] * world
    dist.all_gather_object(out, (W, a, b, cv, ch))
    if rank == 0:
        ret.put(out)
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,port", [("allreduce", 29611), ("sharded_update", 29612), ("state_exchange", 29613)])
def test_sharded_allreduce_equals_single_process(mode, port):
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret, mode)) for r in range(world)]
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
