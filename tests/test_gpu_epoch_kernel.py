"""The persistent multi-mini-batch kernel (csrc/sm100_epoch.cuh, qb_train_steps / qb_*_epoch on small models) against
(1) the same steps run one launch sequence at a time -- already pinned to the oracle by test_gpu_parity.py -- and
(2) the oracle directly at the BASELINE cfg1 shape (784 x 500, CD-1, batch 100) with the Philox stream.
Run with -m gpu."""
import numpy as np
import pytest

from oracle import qarbom_oracle as qo
from tests import philox_ref
from tests.helpers import bf16_round, f32_round, rel_err

pytestmark = pytest.mark.gpu
SEED = 0xE90C


@pytest.fixture(scope="module")
def q():
    import qarbom_b200
    qarbom_b200.build()
    return qarbom_b200


def _model(q, rng, V, H, L, gaussian):
    W = bf16_round(0.05 * rng.standard_normal((V, H)))
    if L:
        r = q.RBMClassifier(V, H, L, W=W, U=bf16_round(0.05 * rng.standard_normal((L, H))))
        r.c[:] = f32_round(0.05 * rng.standard_normal(L))
    else:
        r = (q.GRBM if gaussian else q.RBM)(V, H, W=W)
    r.a[:] = f32_round(0.05 * rng.standard_normal(V))
    r.b[:] = f32_round(0.05 * rng.standard_normal(H))
    return r


CASES = [  # V, H, L, gaussian, B, k, algo, n_batches, n_steps
    (784, 500, 0, False, 100, 1, "CD", 5, 7),      # cfg1 shape; more steps than resident batches: the batch index wraps
    (784, 500, 0, False, 256, 1, "PCD", 3, 5),     # cfg2 shape (chains = batch)
    (70, 33, 0, False, 20, 3, "CD", 4, 6),         # ragged: nothing is a multiple of a tile
    (70, 33, 0, False, 36, 2, "PCD", 3, 4),
    (200, 300, 0, True, 64, 2, "CD", 3, 4),        # Gaussian visibles
    (784, 1000, 10, False, 512, 1, "CD", 2, 3),    # cfg4 shape: label units fused into the visible epilogue
    (112, 64, 6, False, 48, 2, "CD", 3, 4),        # label units in the middle of a 32-lane group
]


@pytest.mark.parametrize("V,H,L,gaussian,B,k,algo,nb,n_steps", CASES)
def test_persistent_kernel_equals_step_by_step(q, V, H, L, gaussian, B, k, algo, nb, n_steps):
    rng = np.random.default_rng(V + H + B)
    X = rng.standard_normal((nb * B, V)).astype(np.float32) if gaussian else (rng.random((nb * B, V)) < 0.3).astype(np.uint8)
    Y = np.eye(L, dtype=np.uint8)[rng.integers(0, L, nb * B)] if L else None
    out = {}
    for mode in (0, 2):
        rbm = _model(q, np.random.default_rng(7), V, H, L, gaussian)
        with q.DeviceRBM(rbm, max_batch=B, precision="bf16", seed=SEED) as dev:
            dev.set_option("epoch_kernel", mode)
            dev.set_data(X, Y)
            if algo == "PCD":
                dev.init_chains(B)
            dev.reset_counters()
            dev.train_steps(algo, 1, n_steps, B, k, 0.05, 0.02)      # starts at batch 1, wraps around
            hist = dev.gemm_histogram()
            assert (sum(n for key, n in hist.items() if key[2] == "epoch") == 1) == (mode == 2), hist
            if mode == 2:
                assert sum(hist.values()) == 1 and dev.counters()[0] <= 2 * nb + 2, (hist, dev.counters())   # (+ lazily built transposed data copies)
            chains = dev.get_chains() if algo == "PCD" else None
            # one more step through the ordinary path: the kernel must leave every piece of state (draw counter, shadows,
            # bias accumulators, chains) exactly where the step-by-step path would
            dev.set_option("epoch_kernel", 0)
            dev.train_steps(algo, 0, 1, B, k, 0.05, 0.02)
            dev.pull_params()
            out[mode] = (rbm.W.copy(), rbm.a.copy(), rbm.b.copy(), None if not L else rbm.U.copy(), None if not L else rbm.c.copy(), chains)
    ref, got = out[0], out[2]
    if algo == "PCD":
        assert np.array_equal(ref[5][0], got[5][0]) and np.array_equal(ref[5][1], got[5][1]), "chain states differ"
    # same GEMMs, same epilogue arithmetic; only the order of the float atomics of the bias sums differs (and feeds back
    # into later probabilities at the 1e-7 level): a few fp32 ulps of the parameter's magnitude (the GRBM case grows |a| to
    # ~1e2 within these steps; two runs of the step-by-step path differ from each other by the same amount)
    def close(r, g):
        return np.max(np.abs(r - g)) <= 1e-5 * max(1.0, float(np.max(np.abs(r))))
    assert close(ref[0], got[0]), np.max(np.abs(ref[0] - got[0]))
    assert close(ref[1], got[1]) and close(ref[2], got[2]), (np.max(np.abs(ref[1] - got[1])), np.max(np.abs(ref[2] - got[2])))
    if L:
        assert close(ref[3], got[3]) and close(ref[4], got[4])


def test_cfg1_epoch_kernel_matches_oracle(q):
    """BASELINE configs[0]: Binary RBM 784 x 500, CD-1, batch 100 -- five consecutive mini-batch steps in one launch against
    the oracle's batched formulation (tested equal to the per-sample restatement of src/training/train_cd.jl:2-21) fed the
    Philox numbers of the stream contract, re-synchronised with the bf16-rounded weights after every step."""
    V, H, B, nb, lr = 784, 500, 100, 5, 0.05
    rng = np.random.default_rng(1)
    X = (rng.random((nb * B, V)) < 0.1307).astype(np.uint8)
    m = qo.make_model(V, H, W=bf16_round(0.01 * rng.standard_normal((V, H))))
    W_init = m.W.copy()
    rbm = q.RBM(V, H, W=W_init)
    Ws = []
    with q.DeviceRBM(rbm, max_batch=B, precision="bf16", seed=SEED) as dev:
        dev.set_data(X)
        dev.set_option("epoch_kernel", 2)
        for t in range(nb):          # one call per step so that the oracle can follow the device's bf16 rounding step by step ...
            dev.train_steps("CD", t, 1, B, 1, lr)
            dev.pull_params()
            Ws.append((rbm.W.copy(), rbm.a.copy(), rbm.b.copy()))
    Xf = X.astype(np.float64)
    for t in range(nb):
        W0, a0, b0 = m.W.copy(), m.a.copy(), m.b.copy()
        Wd0 = W0 if t == 0 else Ws[t - 1][0]
        qo.batched_step(m, Xf[t * B:(t + 1) * B], None, "CD", 1, lr, U_h=philox_ref.uniforms(SEED, 2 * t, B, H)[None],
                        U_v=philox_ref.uniforms(SEED, 2 * t + 1, B, V)[None])
        assert rel_err(Ws[t][0] - Wd0, m.W - W0) < 2e-2, f"step {t}: dW rel err {rel_err(Ws[t][0] - Wd0, m.W - W0):.3e}"
        assert rel_err(Ws[t][1] - a0, m.a - a0) < 2e-2 and rel_err(Ws[t][2] - b0, m.b - b0) < 2e-2
        m.W[...] = bf16_round(Ws[t][0]); m.a[...] = Ws[t][1]; m.b[...] = Ws[t][2]
    # ... and all five steps in ONE launch give the same parameters
    rbm2 = q.RBM(V, H, W=W_init)
    with q.DeviceRBM(rbm2, max_batch=B, precision="bf16", seed=SEED) as dev:
        dev.set_data(X)
        dev.set_option("epoch_kernel", 2)
        dev.reset_counters()
        dev.train_steps("CD", 0, nb, B, 1, lr)
        assert dev.gemm_histogram().get(("1cta", 32, "epoch")) == 1
        dev.pull_params()
    assert np.max(np.abs(rbm2.W - Ws[-1][0])) < 1e-5   # (float atomics of the bias sums differ in order: a few fp32 ulps per step)


def test_epoch_call_reports_phase_times_and_equals_steps(q):
    V, H, B, nb = 784, 500, 100, 12
    rng = np.random.default_rng(3)
    X = (rng.random((nb * B, V)) < 0.13).astype(np.uint8)
    res = []
    for mode in ("epoch", "steps"):
        rbm = _model(q, np.random.default_rng(5), V, H, 0, False)
        with q.DeviceRBM(rbm, max_batch=B, precision="bf16", seed=SEED) as dev:
            dev.set_data(X)
            if mode == "epoch":
                dev.reset_counters()
                t = dev.cd_epoch(B, 1, 0.05)
                assert all(x >= 0 for x in t) and sum(t) > 0
                assert sum(n for key, n in dev.gemm_histogram().items() if key[2] == "epoch") == 1
            else:
                dev.set_option("epoch_kernel", 0)
                for i in range(nb):
                    dev.cd_step(i, B, 1, 0.05)
            dev.pull_params()
            res.append(rbm.W.copy())
    assert np.max(np.abs(res[0] - res[1])) < 1e-5   # 12 steps; a flipped sample would show as lr / B = 5e-4
