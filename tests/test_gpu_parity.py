"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Run with -m gpu on a B200.

Validation mode (fp32 kernels, injected streams): sampled states must match the Float64
oracle bit-exactly (streams are guard-banded by 1e-4 so an fp32 evaluation of p cannot flip
a draw) and W/a/b must agree within 1e-5 relative.
bf16 mode (tcgen05 path): the oracle runs on the bf16-rounded weights the tensor cores see
and is re-synchronised with the device's fp32 masters after every step; states must again
match exactly, parameter *increments* within the looser 2e-2 relative (the gradient GEMM
consumes bf16-rounded probabilities).
"""
import numpy as np
import pytest

from oracle import qarbom_oracle as qo
from tests import philox_ref
from tests.helpers import (DenseStreams, bf16_round, chains_from_arrays, chains_to_arrays, f32_round, make_oracle_model,
                           rel_err, to_host_rbm, u24)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q():
    import qarbom_b200
    qarbom_b200.build()
    return qarbom_b200


# "fp32x3" = QB_PREC_FP32X3, the validation mode ON the tensor cores (three bf16 components per fp32 operand): fp32 tolerances
TOL = {"fp32": dict(w=1e-5, dw=2e-4, p=2e-6), "fp32x3": dict(w=1e-5, dw=2e-4, p=2e-6), "bf16": dict(w=None, dw=2e-2, p=2e-6)}


def _resync(m, dev, prec):
    """bf16 mode: continue the oracle from the device's masters (weights as the tensor cores see them)."""
    dev.pull_params()
    r = dev.rbm
    m.W[...] = bf16_round(r.W) if prec == "bf16" else r.W
    m.a[...] = r.a
    m.b[...] = r.b
    if m.is_classifier:
        m.U[...] = bf16_round(r.U) if prec == "bf16" else r.U
        m.c[...] = r.c


def _data(rng, n, V, L, gaussian, bf16):
    X = rng.standard_normal((n, V)) if gaussian else rng.integers(0, 2, (n, V)).astype(np.float64)
    X = bf16_round(X) if bf16 else f32_round(X)
    Y = np.eye(L)[rng.integers(0, L, n)] if L else None
    return X, Y


# ------------------------------------------------------------------------------ conditionals
@pytest.mark.parametrize("prec", ["fp32", "fp32x3", "bf16"])
@pytest.mark.parametrize("V,H,L", [(3, 2, 0), (16, 8, 0), (70, 33, 5), (784, 500, 10)])
def test_conditional_prob_h_and_reconstruct(q, prec, V, H, L):
    rng = np.random.default_rng(V * 1000 + H)
    m = make_oracle_model(rng, V, H, L, bf16=(prec == "bf16"), scale=0.2)
    X, Y = _data(rng, 37, V, L, False, prec == "bf16")
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=64, precision=prec) as dev:
        ph = dev.conditional_prob_h(X)
        ref = np.stack([qo.conditional_prob_h(m, x) for x in X])
        assert np.max(np.abs(ph - ref)) < TOL[prec]["p"]
        if L:
            ph3 = dev.conditional_prob_h(X, Y)
            ref3 = np.stack([qo.conditional_prob_h(m, x, y) for x, y in zip(X, Y)])
            assert np.max(np.abs(ph3 - ref3)) < TOL[prec]["p"]
        rec = dev.reconstruct(X)
        if prec == "bf16":  # hidden probabilities are rounded to bf16 before the second GEMM
            refr = np.stack([qo.conditional_prob_v(m, bf16_round(qo.conditional_prob_h(m, x))) for x in X])
        else:
            refr = np.stack([qo.reconstruct(m, x) for x in X])
        assert np.max(np.abs(rec - refr)) < 5e-6


# ------------------------------------------------------------------------------ PCD
PCD_CASES = [  # V, H, L, gaussian, B, k, steps
    (3, 2, 0, False, 2, 1, 3),        # the reference's own test scenario shape (test/generative.jl:22-38)
    (16, 8, 0, False, 8, 2, 3),
    (64, 32, 0, False, 16, 3, 2),
    (70, 33, 0, False, 13, 2, 2),     # ragged: nothing is a multiple of the tile sizes
    (784, 500, 0, False, 64, 1, 2),   # cfg-shaped
    (3, 2, 1, False, 2, 1, 3),        # test/classification.jl shape
    (40, 24, 5, False, 12, 2, 3),
    (24, 16, 0, True, 8, 2, 2),       # GRBM
    (24, 16, 3, True, 8, 1, 2),       # GRBMClassifier
]


@pytest.mark.parametrize("prec", ["fp32", "fp32x3", "bf16"])
@pytest.mark.parametrize("V,H,L,gaussian,B,k,steps", PCD_CASES)
def test_pcd_steps_match_oracle(q, prec, V, H, L, gaussian, B, k, steps):
    bf = prec == "bf16"
    rng = np.random.default_rng(7 + V + 13 * H + L)
    m = make_oracle_model(rng, V, H, L, gaussian, bf16=bf)
    if bf and gaussian:
        m.v_store = bf16_round   # the device keeps real-valued visible states in bf16: the oracle rounds them where they are stored
    X, Y = _data(rng, steps * B, V, L, gaussian, bf)
    cv0, ch0 = u24(rng, B, V), u24(rng, B, H)
    cy0 = u24(rng, B, L) if L else None
    if bf:
        cv0, ch0 = bf16_round(cv0), bf16_round(ch0)
        cy0 = bf16_round(cy0) if L else None
    chains = chains_from_arrays(cv0, ch0, cy0)
    lr, llr = 0.05, 0.02
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision=prec) as dev:
        dev.set_data(X, Y)
        dev.init_chains(B, cv0, ch0, cy0)
        for t in range(steps):
            ds = DenseStreams(rng, k, B, V, H, L, gaussian)
            su, sz = ds.serial("pcd")
            st = qo.GuardedStream(su, sz)
            W0, a0, b0 = m.W.copy(), m.a.copy(), m.b.copy()
            qo.persistent_contrastive_divergence(m, X, [range(t * B, (t + 1) * B)], chains, steps=k, learning_rate=lr,
                                                 stream=st, y=Y, label_learning_rate=llr)
            assert st.upos == su.size
            ds.absorb(st.u, "pcd")
            dev.rbm.W[...] = 0  # make sure pull_params really reads the device
            dev.inject_streams(ds.U_h, ds.U_v, ds.U_y, ds.Z_v)
            dev.pcd_step(t, B, k, lr, llr)
            dv, dh, dy = dev.get_chains()
            ov, oh, oy = chains_to_arrays(chains)
            assert np.array_equal(dh, oh), f"hidden chain states differ at step {t}"
            if gaussian and bf:
                # bf16(fp32 mean + z) against bf16(Float64 mean + z): equal except where the two means straddle a rounding
                # boundary -- then one bf16 ulp (2^-8 relative) apart
                if L:   # classifier PCD advances the chains AFTER the update (train_pcd.jl:89-92): the device's chains see
                        # W_{t+1} rounded to bf16, the oracle (re-synchronised once per step) the unrounded one
                    assert np.all(np.abs(dv - ov) <= 2.0 ** -6 * np.abs(ov) + 2e-2)
                else:
                    assert np.all(np.abs(dv - ov) <= 2.0 ** -7 * np.abs(ov) + 1e-6) and np.mean(dv == ov) > 0.98
            elif gaussian:
                assert np.max(np.abs(dv - ov)) < 1e-5
            else:
                assert np.array_equal(dv, ov), f"visible chain states differ at step {t}"
            if L:
                assert np.array_equal(dy, oy), f"label chain states differ at step {t}"
            dev.pull_params()
            r = dev.rbm
            if bf:
                # increments (device master is fp32; oracle started from the same rounded weights)
                assert rel_err(r.a - a0, m.a - a0) < TOL[prec]["dw"] or np.linalg.norm(m.a - a0) < 1e-12
                dW_o = m.W - W0
                Wdev_before = dev_prev_W if t > 0 else W0
                assert rel_err(r.W - Wdev_before, dW_o) < TOL[prec]["dw"]
                dev_prev_W = r.W.copy()
                _resync(m, dev, prec)
            else:
                assert rel_err(r.W, m.W) < TOL[prec]["w"]
                assert rel_err(r.W - W0, m.W - W0) < TOL[prec]["dw"]
                assert np.max(np.abs(r.a - m.a)) < 1e-6 and np.max(np.abs(r.b - m.b)) < 1e-6
                if L:
                    assert rel_err(r.U, m.U) < TOL[prec]["w"] and np.max(np.abs(r.c - m.c)) < 1e-6
                dev_prev_W = r.W.copy()


# ------------------------------------------------------------------------------ CD
CD_CASES = [  # V, H, L, gaussian, B, k, n_batches
    (3, 2, 0, False, 1, 3, 6),         # reference scenario: online CD-3 (test/generative.jl:4-20)
    (16, 8, 0, False, 1, 2, 5),
    (64, 32, 0, False, 16, 2, 2),      # mini-batch extension
    (784, 500, 0, False, 32, 1, 2),    # cfg1-shaped
    (3, 2, 1, False, 1, 3, 6),
    (50, 20, 10, False, 16, 1, 2),     # cfg4-shaped (small)
    (24, 16, 0, True, 8, 3, 2),        # GRBM CD-3
]


@pytest.mark.parametrize("prec", ["fp32", "fp32x3", "bf16"])
@pytest.mark.parametrize("V,H,L,gaussian,B,k,nb", CD_CASES)
def test_cd_steps_match_oracle(q, prec, V, H, L, gaussian, B, k, nb):
    bf = prec == "bf16"
    rng = np.random.default_rng(11 + V + 13 * H + L + B)
    m = make_oracle_model(rng, V, H, L, gaussian, bf16=bf)
    if bf and gaussian:
        m.v_store = bf16_round   # (see test_pcd_steps_match_oracle)
    X, Y = _data(rng, nb * B, V, L, gaussian, bf)
    lr, llr = 0.05, 0.02
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision=prec) as dev:
        dev.set_data(X, Y)
        for t in range(nb):
            ds = DenseStreams(rng, k, B, V, H, L, gaussian)
            su, sz = ds.serial("cd")
            st = qo.GuardedStream(su, sz)
            W0 = m.W.copy()
            tr = []
            sl = slice(t * B, (t + 1) * B)
            qo.contrastive_divergence(m, X[sl], steps=k, learning_rate=lr, stream=st, y=None if Y is None else Y[sl],
                                      label_learning_rate=llr, batch_size=B, trace=tr)
            ds.absorb(st.u, "cd")
            dev.inject_streams(ds.U_h, ds.U_v, ds.U_y, ds.Z_v)
            Wd0 = dev.rbm.W.copy() if t == 0 else Wd_prev
            dev.cd_step(t, B, k, lr, llr)
            dev.pull_params()
            r = dev.rbm
            if bf:
                assert rel_err(r.W - Wd0, m.W - W0) < TOL[prec]["dw"]
                Wd_prev = r.W.copy()
                _resync(m, dev, prec)
            else:
                assert rel_err(r.W, m.W) < TOL[prec]["w"]
                assert rel_err(r.W - W0, m.W - W0) < TOL[prec]["dw"]
                assert np.max(np.abs(r.a - m.a)) < 1e-6 and np.max(np.abs(r.b - m.b)) < 1e-6
                if L:
                    assert rel_err(r.U, m.U) < TOL[prec]["w"]
                Wd_prev = r.W.copy()


# ------------------------------------------------------------------------------ bf16 Gaussian visibles (tolerance)
def test_bf16_grbm_single_step_tolerance(q):
    rng = np.random.default_rng(99)
    V, H, B, k = 48, 32, 16, 1
    m0 = make_oracle_model(rng, V, H, 0, True, bf16=True)
    m = qo.copy_rbm(m0)
    X, _ = _data(rng, B, V, 0, True, True)
    ds = DenseStreams(rng, k, B, V, H, 0, True)
    su, sz = ds.serial("cd")
    st = qo.GuardedStream(su, sz)
    qo.contrastive_divergence(m, X, steps=k, learning_rate=0.01, stream=st, batch_size=B)
    ds.absorb(st.u, "cd")
    with q.DeviceRBM(to_host_rbm(q, m0), max_batch=B, precision="bf16") as dev:
        dev.set_data(X)
        dev.inject_streams(ds.U_h, None, None, ds.Z_v)
        dev.cd_step(0, B, k, 0.01)
        dev.pull_params()
        # visible samples are rounded to bf16 on the device: 2^-8 relative on v -> looser tolerance on dW
        assert rel_err(dev.rbm.W - m0.W, m.W - m0.W) < 5e-2


# ------------------------------------------------------------------------------ production RNG (Philox)
@pytest.mark.parametrize("prec", ["fp32", "fp32x3", "bf16"])
def test_philox_mode_matches_oracle_fed_with_the_same_stream(q, prec):
    """Device Philox == numpy Philox: the oracle, fed the uniforms the stream contract
    defines, must reproduce the device's chain states (outside a 1e-4 guard band)."""
    bf = prec == "bf16"
    rng = np.random.default_rng(5)
    V, H, B, k, seed = 96, 40, 24, 2, 0xC0FFEE1234
    m = make_oracle_model(rng, V, H, bf16=bf)
    X, _ = _data(rng, B, V, 0, False, bf)
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision=prec, seed=seed) as dev:
        dev.set_data(X)
        dev.init_chains(B)  # draws 0 (v) and 1 (h)
        cv0, ch0, _ = dev.get_chains()
        ev, eh = philox_ref.uniforms(seed, 0, B, V), philox_ref.uniforms(seed, 1, B, H)
        if bf:
            ev, eh = bf16_round(ev), bf16_round(eh)
        assert np.array_equal(cv0, ev) and np.array_equal(ch0, eh)
        dev.pcd_step(0, B, k, 0.05)
        dv, dh, _ = dev.get_chains()
    chains = chains_from_arrays(cv0, ch0)
    U_h = np.stack([philox_ref.uniforms(seed, 2 + 2 * s, B, H) for s in range(k)])
    U_v = np.stack([philox_ref.uniforms(seed, 3 + 2 * s, B, V) for s in range(k)])
    # replay step by step so that near-ties can be excluded from the comparison
    near = 0
    for s in range(k):
        for i, c in enumerate(chains):
            p = qo.conditional_prob_h(m, c.v)
            c.h = (U_h[s, i] < p).astype(np.float64)
            near += int(np.sum(np.abs(U_h[s, i] - p) < 1e-5))
            pv = qo.conditional_prob_v(m, c.h)
            c.v = (U_v[s, i] < pv).astype(np.float64)
            near += int(np.sum(np.abs(U_v[s, i] - pv) < 1e-5))
    ov, oh, _ = chains_to_arrays(chains)
    if near == 0:
        assert np.array_equal(dv, ov) and np.array_equal(dh, oh)
    else:  # extremely unlikely with these sizes; still require near-total agreement
        assert np.mean(dv != ov) < 1e-3 and np.mean(dh != oh) < 1e-3


def test_philox_bernoulli_frequencies(q):
    """Statistical check of fused sampling: with W = 0 the hidden units fire with sigma(b)."""
    V, H, B = 32, 64, 4096
    m = qo.make_model(V, H, W=np.zeros((V, H)))
    m.b[:] = np.linspace(-2, 2, H)
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision="bf16", seed=77) as dev:
        dev.set_data(np.zeros((B, V)))
        dev.init_chains(B)
        dev.pcd_step(0, B, 1, 0.0)
        _, h, _ = dev.get_chains()
    p = 1 / (1 + np.exp(-m.b))
    z = (h.mean(0) - p) / np.sqrt(p * (1 - p) / B)
    assert np.max(np.abs(z)) < 4.5


# ------------------------------------------------------------------------------ evaluation
@pytest.mark.parametrize("prec", ["fp32", "fp32x3", "bf16"])
@pytest.mark.parametrize("V,H,L,gaussian", [(3, 2, 0, False), (70, 33, 0, False), (70, 33, 4, False), (40, 24, 0, True)])
def test_eval_mse_and_free_energy(q, prec, V, H, L, gaussian):
    bf = prec == "bf16"
    rng = np.random.default_rng(3 + V)
    m = make_oracle_model(rng, V, H, L, gaussian, bf16=bf, scale=0.2)
    n = 150  # > max_batch: exercises chunking
    X, Y = _data(rng, n, V, L, gaussian, bf)
    Z = f32_round(rng.standard_normal((n, V))) if gaussian else None
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=64, precision=prec) as dev:
        dev.set_data(X, Y)
        mse_res = dev.eval_mse(None, Z)
        mse_host = dev.eval_mse(X, Z)
        fe = dev.eval_free_energy()
    ref_mse = qo.evaluate_mse(m, X, qo.Stream(normals=Z))
    tol = 2e-3 if bf else 1e-5
    assert abs(mse_res - ref_mse) / ref_mse < tol and abs(mse_host - ref_mse) / ref_mse < tol
    ref_fe = qo.mean_free_energy(m, X, Y)
    assert abs(fe - ref_fe) / abs(ref_fe) < (1e-3 if bf else 1e-5)


# ------------------------------------------------------------------------------ epoch driver / train!
def test_epoch_equals_steps_and_drops_remainder(q):
    rng = np.random.default_rng(8)
    V, H, B, k, n = 20, 12, 8, 1, 27  # 3 batches + 3 dropped rows (src/utils.jl:13-26)
    m = make_oracle_model(rng, V, H)
    X, _ = _data(rng, n, V, 0, False, False)
    out = []
    for mode in ("epoch", "steps"):
        with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision="fp32", seed=9) as dev:
            dev.set_data(X)
            dev.init_chains(B)
            if mode == "epoch":
                t = dev.pcd_epoch(B, k, 0.05)
                assert all(x >= 0 for x in t) and sum(t) > 0
            else:
                for i in range(n // B):
                    dev.pcd_step(i, B, k, 0.05)
            dev.pull_params()
            out.append(dev.rbm.W.copy())
    assert np.max(np.abs(out[0] - out[1])) < 1e-6  # identical up to the order of the float atomics in the bias sums


@pytest.mark.parametrize("method", ["CD", "PCD"])
def test_train_reference_scenario(q, method, tmp_path):
    """test/generative.jl:4-38: RBM(3,2) on [1,0,0] x 1000, lr = 0.01/j^0.8, early stopping.
    The reference asserts nothing; we assert what its docs promise: unit 1 is reconstructed."""
    x = [[1, 0, 0] for _ in range(1000)]
    rbm = q.RBM(3, 2, rng=np.random.default_rng(1))
    kw = dict(n_epochs=20, learning_rate=[0.05 / (j ** 0.8) for j in range(1, 1001)], early_stopping=True,
              file_path=str(tmp_path / "m.csv"), precision="fp32", verbose=False)
    if method == "CD":
        hist = q.train(rbm, x, q.CD, gibbs_steps=3, **kw)
    else:
        hist = q.train(rbm, x, q.PCD, batch_size=2, **kw)
    rec = q.reconstruct(rbm, [1, 0, 0])
    assert rec[0] > 0.5 > max(rec[1], rec[2])
    assert len(hist["mse"]) >= 2 and min(hist["mse"]) <= hist["mse"][0]
    assert (tmp_path / "m.csv").read_text().splitlines()[0] == "mse"


def test_argument_errors(q):
    rbm = q.RBM(4, 3)
    with q.DeviceRBM(rbm, max_batch=4, precision="fp32") as dev:
        with pytest.raises(q.QbError):
            dev.pcd_step(0, 4, 1, 0.1)          # no data
        dev.set_data(np.zeros((8, 4)))
        with pytest.raises(q.QbError):
            dev.pcd_step(0, 4, 1, 0.1)          # no chains
        dev.init_chains(4)
        with pytest.raises(q.QbError):
            dev.pcd_step(2, 4, 1, 0.1)          # batch index out of range
        with pytest.raises(q.QbError):
            dev.pcd_step(0, 8, 1, 0.1)          # B > max_batch
        dev.pcd_step(1, 4, 1, 0.1)
        with pytest.raises(q.QbError):
            dev.cd_step(0, 4, 0, 0.1)           # CD needs k >= 1 (k = 0 is the qb_pcd_step-only "external negative samples" form)
        with pytest.raises(q.QbError):
            dev.set_data(np.zeros((0, 4)))      # empty dataset
        with pytest.raises(q.QbError):
            dev.eval_mse(np.zeros((0, 4)))      # empty evaluation set
        with pytest.raises(q.QbError):
            dev.set_option("no_such_option", 1)
        with pytest.raises(q.QbError):
            dev.set_option("sharded_update", 1)  # needs bf16 and a communicator world > 1
        with pytest.raises(q.QbError):
            dev.fast_pcd_epoch(2, 1, 0.1, 0.1)   # FastPCD pairs sample i with chain i: chains != batch rows
        assert dev.eval_mse() >= 0.0             # the failed calls left the handle usable
        dev.set_data(np.zeros((2, 4)))           # fewer samples than one mini-batch: `_set_mini_batches` drops them all,
        dev.pull_params(); W0 = rbm.W.copy()     # the epoch trains nothing (src/utils.jl:13-26)
        assert dev.pcd_epoch(4, 1, 0.1) == (0.0, 0.0, 0.0) and dev.cd_epoch(4, 1, 0.1) == (0.0, 0.0, 0.0)
        dev.pull_params()
        assert np.array_equal(rbm.W, W0)
    with pytest.raises(q.QbError):
        q.DeviceRBM(q.RBM(4, 3), max_batch=0)    # invalid configuration


# ------------------------------------------------------------------------------ optimizer extension
@pytest.mark.parametrize("V,H,L", [(40, 24, 0), (130, 70, 5)])
def test_momentum_and_weight_decay_extension(q, V, H, L):
    """momentum / weight decay are extensions (absent from update_rbm!, src/rbm.jl:152-163); the
    oracle defines them (apply_update) and momentum = wd = 0 is covered by every other test."""
    rng = np.random.default_rng(21 + V)
    B, k, steps = 16, 1, 4
    m = make_oracle_model(rng, V, H, L)
    X, Y = _data(rng, steps * B, V, L, False, False)
    cv0, ch0 = u24(rng, B, V), u24(rng, B, H)
    cy0 = u24(rng, B, L) if L else None
    chains = chains_from_arrays(cv0, ch0, cy0)
    opt = qo.OptState()
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision="fp32") as dev:
        dev.set_optimizer(0.5, 1e-3)
        dev.set_data(X, Y)
        dev.init_chains(B, cv0, ch0, cy0)
        for t in range(steps):
            ds = DenseStreams(rng, k, B, V, H, L, False)
            su, sz = ds.serial("pcd")
            st = qo.GuardedStream(su, sz)
            qo.persistent_contrastive_divergence(m, X, [range(t * B, (t + 1) * B)], chains, steps=k, learning_rate=0.05,
                                                 stream=st, y=Y, label_learning_rate=0.02, momentum=0.5, weight_decay=1e-3,
                                                 opt=opt)
            ds.absorb(st.u, "pcd")
            dev.inject_streams(ds.U_h, ds.U_v, ds.U_y, ds.Z_v)
            dev.pcd_step(t, B, k, 0.05, 0.02)
        dev.pull_params()
        assert rel_err(dev.rbm.W, m.W) < 1e-5 and np.max(np.abs(dev.rbm.b - m.b)) < 1e-6
        if L:
            assert rel_err(dev.rbm.U, m.U) < 1e-5 and np.max(np.abs(dev.rbm.c - m.c)) < 1e-6


# ------------------------------------------------------------------------------ 2-CTA (cta_group::2) kernel variant
@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("V,H,B", [(300, 260, 200), (512, 512, 256), (70, 33, 13)])
def test_2cta_variant_matches_oracle(q, bn, V, H, B):
    """Forces the cta_group::2 kernel (normally used only when 256-row tiles fill the machine) on
    small and ragged shapes: injected streams (generic epilogue) and Philox (fast epilogue)."""
    rng = np.random.default_rng(1234 + V + bn)
    k = 2
    m = make_oracle_model(rng, V, H, bf16=True)
    X, _ = _data(rng, B, V, 0, False, True)
    cv0, ch0 = bf16_round(u24(rng, B, V)), bf16_round(u24(rng, B, H))
    chains = chains_from_arrays(cv0, ch0)
    ds = DenseStreams(rng, k, B, V, H)
    st = qo.GuardedStream(ds.serial("pcd")[0])
    m_o = qo.copy_rbm(m)
    qo.persistent_contrastive_divergence(m_o, X, [range(0, B)], chains, steps=k, learning_rate=0.05, stream=st)
    ds.absorb(st.u, "pcd")
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision="bf16", seed=5) as dev:
        dev.set_option("use_2cta", 2)
        dev.set_option("force_bn", bn)
        dev.set_data(X)
        dev.init_chains(B, cv0, ch0)
        dev.inject_streams(ds.U_h, ds.U_v)
        dev.pcd_step(0, B, k, 0.05)
        dv, dh, _ = dev.get_chains()
        ov, oh, _ = chains_to_arrays(chains)
        assert np.array_equal(dh, oh) and np.array_equal(dv, ov)
        dev.pull_params()
        assert rel_err(dev.rbm.W - m.W, m_o.W - m.W) < 2e-2
    # Philox / fast epilogue: 2-CTA result must equal the 1-CTA result
    outs = []
    for two in (0, 2):
        with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision="bf16", seed=5) as dev:
            dev.set_option("use_2cta", two)
            dev.set_option("force_bn", bn)
            dev.set_data(X)
            dev.init_chains(B, cv0, ch0)
            dev.pcd_step(0, B, k, 0.05)
            v, h, _ = dev.get_chains()
            dev.pull_params()
            outs.append((v, h, dev.rbm.W.copy(), dev.rbm.a.copy(), dev.rbm.b.copy()))
    for a, b in zip(outs[0][:3], outs[1][:3]):  # states and W (same MMA order) are bit-identical
        assert np.array_equal(a, b)
    for a, b in zip(outs[0][3:], outs[1][3:]):  # bias sums are float atomics: order-dependent last bits
        assert np.max(np.abs(a - b)) < 1e-6


# ------------------------------------------------------------------------------ stale-chains (overlapped allreduce) mode
def test_stale_chains_mode_matches_its_definition(q):
    """Declared extension (SURVEY section 8(e) option 2): the chains of step t advance with the weights
    of step t-1 while the allreduce + update of step t-1 are still in flight; the positive phase uses
    the latest weights.  Emulated here in numpy (batched form, bf16-rounded weight versions)."""
    rng = np.random.default_rng(77)
    V, H, B, k, steps, lr = 96, 64, 32, 2, 4, 0.05
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    m = make_oracle_model(rng, V, H, bf16=True)
    X, _ = _data(rng, steps * B, V, 0, False, True)
    cv, ch = bf16_round(u24(rng, B, V)), bf16_round(u24(rng, B, H))
    Wm, a, b = m.W.copy(), m.a.copy(), m.b.copy()          # fp32 masters
    versions = [(bf16_round(Wm), a.copy(), b.copy())]        # shadow versions as the tensor cores see them

    def bern(u, p):  # guard-band near ties in place (same trick as GuardedStream)
        near = np.abs(u - p) < 1e-4
        out = u < p
        u[near] = np.where(out[near], p[near] - 2e-4, p[near] + 2e-4).astype(np.float32)
        return out.astype(np.float64)

    with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision="bf16") as dev:
        dev.set_option("stale_chains", 1)
        dev.set_data(X)
        dev.init_chains(B, cv, ch)
        for t in range(steps):
            Wc, ac, bc = versions[max(t - 1, 0)]              # chains: one update old
            Wp, ap, bp = versions[t]                          # positive phase: latest
            U_h, U_v = u24(rng, k, B, H), u24(rng, k, B, V)
            for s in range(k):
                ch = bern(U_h[s], sig(bc + cv @ Wc))
                cv = bern(U_v[s], sig(ac + ch @ Wc.T))
            xb = X[t * B:(t + 1) * B]
            hd = sig(bp + xb @ Wp)
            Wm = Wm + (lr / B) * (xb.T @ hd - cv.T @ ch)
            a = a + (lr / B) * (xb - cv).sum(0)
            b = b + (lr / B) * (hd - ch).sum(0)
            versions.append((bf16_round(Wm), f32_round(a), f32_round(b)))
            dev.inject_streams(U_h, U_v)
            dev.pcd_step(t, B, k, lr)
            dv, dh, _ = dev.get_chains()
            assert np.array_equal(dh, ch) and np.array_equal(dv, cv), f"chain states differ at step {t}"
            dev.pull_params()
            # keep the emulation on the device's masters so rounding differences cannot accumulate
            assert rel_err(dev.rbm.W - versions[t][0], Wm - versions[t][0], atol=1e-6) < 5e-2 or t == 0
            Wm, a, b = dev.rbm.W.copy(), dev.rbm.a.copy(), dev.rbm.b.copy()
            versions[-1] = (bf16_round(Wm), a.copy(), b.copy())
        # leaving the mode makes the last update current
        dev.set_option("stale_chains", 0)
        ph = dev.conditional_prob_h(X[:4])
        assert np.max(np.abs(ph - sig(b + X[:4] @ bf16_round(Wm)))) < 2e-6


# ------------------------------------------------------------------------------ classifier evaluation
@pytest.mark.parametrize("prec", ["fp32", "fp32x3", "bf16"])
def test_classify_and_accuracy(q, prec):
    bf = prec == "bf16"
    rng = np.random.default_rng(17)
    V, H, L, n = 70, 33, 6, 90
    m = make_oracle_model(rng, V, H, L, bf16=bf)
    X, Y = _data(rng, n, V, L, False, bf)
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=32, precision=prec) as dev:
        dev.set_data(X, Y)
        p_res = dev.conditional_prob_y_given_v()
        p_host = dev.conditional_prob_y_given_v(X)
        acc = dev.eval_accuracy(None, Y)
    if bf:
        ref = np.stack([qo.conditional_prob_y_given_h(m, bf16_round(qo.conditional_prob_h(m, x))) for x in X])
    else:
        ref = np.stack([qo.conditional_prob_y_given_v(m, x) for x in X])
    assert np.max(np.abs(p_res - ref)) < 5e-6 and np.max(np.abs(p_host - ref)) < 5e-6
    assert abs(acc - qo.evaluate_accuracy(m, X, Y)) < 1e-12


def test_train_classifier_reference_scenario(q, tmp_path):
    """test/classification.jl:4-44: RBMClassifier(3,2,1) on x=[1,0,0], y=[1]; default metric Accuracy."""
    x = [[1, 0, 0] for _ in range(200)]
    y = [[1] for _ in range(200)]
    rbm = q.RBMClassifier(3, 2, 1, rng=np.random.default_rng(2))
    lr = [0.01 / (j ** 0.8) for j in range(1, 1001)]
    hist = q.train(rbm, x, q.PCD, y, n_epochs=5, batch_size=2, learning_rate=lr, label_learning_rate=lr, early_stopping=True,
                   file_path=str(tmp_path / "c.csv"), precision="fp32", verbose=False)
    assert list(hist) == ["accuracy"] and len(hist["accuracy"]) >= 2
    assert hist["accuracy"][-1] == 1.0   # a single label unit: softmax over one unit is always 1
    assert (tmp_path / "c.csv").read_text().splitlines()[0] == "accuracy"


# ------------------------------------------------------------------------------ host-batch (end-to-end) path
@pytest.mark.parametrize("prec,L", [("fp32", 0), ("fp32x3", 0), ("fp32x3", 4), ("bf16", 0), ("bf16", 4)])
def test_step_host_equals_resident_step_with_and_without_prefetch(q, prec, L):
    rng = np.random.default_rng(31)
    V, H, B, k, steps = 72, 40, 24, 2, 4
    m = make_oracle_model(rng, V, H, L, bf16=(prec == "bf16"))
    X, Y = _data(rng, steps * B, V, L, False, prec == "bf16")
    Xu = X.astype(np.uint8)
    Yu = None if Y is None else Y.astype(np.uint8)
    results = []
    for mode in ("resident", "host", "host+prefetch"):
        rbm = to_host_rbm(q, m)
        with q.DeviceRBM(rbm, max_batch=B, precision=prec, seed=11) as dev:
            dev.init_chains(B)
            if mode == "resident":
                dev.set_data(Xu, Yu)
            for t in range(steps):
                sl = slice(t * B, (t + 1) * B)
                if mode == "resident":
                    dev.pcd_step(t, B, k, 0.05, 0.02)
                else:
                    nxt = slice((t + 1) * B, (t + 2) * B) if (mode == "host+prefetch" and t + 1 < steps) else None
                    err = dev.step_host("PCD", Xu[sl], None if Yu is None else Yu[sl], k, 0.05, 0.02,
                                        next_x=None if nxt is None else Xu[nxt],
                                        next_y=None if (nxt is None or Yu is None) else Yu[nxt])
                    assert np.isfinite(err) and err >= 0
            dev.pull_params()
            v, h, _ = dev.get_chains()
            results.append((rbm.W.copy(), v, h))
    for W, v, h in results[1:]:
        assert np.array_equal(v, results[0][1]) and np.array_equal(h, results[0][2])
        # bias sums are float atomics (order-dependent last bits) and feed back into the probabilities
        assert np.max(np.abs(W - results[0][0])) < 1e-6


# ------------------------------------------------------------------------------ more chains than rows (extension)
@pytest.mark.parametrize("prec", ["fp32", "fp32x3", "bf16"])
@pytest.mark.parametrize("V,H,L,B,Nc", [(40, 24, 0, 8, 20), (40, 24, 0, 12, 4), (30, 18, 3, 8, 16)])
def test_pcd_with_n_chains_different_from_batch(q, prec, V, H, L, B, Nc):
    """BASELINE.json config 2 ("1000 persistent chains, batch 256"): n_chains != batch_size is a declared
    extension (the reference ties them, train_pcd.jl:162); the oracle defines it, n_chains == batch is the
    reference and is what every other PCD test covers."""
    bf = prec == "bf16"
    rng = np.random.default_rng(55 + V + Nc)
    k, steps, lr, llr = 2, 3, 0.05, 0.02
    m = make_oracle_model(rng, V, H, L, bf16=bf)
    X, Y = _data(rng, steps * B, V, L, False, bf)
    cv0, ch0 = u24(rng, Nc, V), u24(rng, Nc, H)
    cy0 = u24(rng, Nc, L) if L else None
    if bf:
        cv0, ch0 = bf16_round(cv0), bf16_round(ch0)
        cy0 = bf16_round(cy0) if L else None
    chains = chains_from_arrays(cv0, ch0, cy0)
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=max(B, Nc), precision=prec) as dev:
        dev.set_data(X, Y)
        dev.init_chains(Nc, cv0, ch0, cy0)
        Wprev = m.W.copy()
        for t in range(steps):
            ds = DenseStreams(rng, k, Nc, V, H, L, False)
            st = qo.GuardedStream(ds.serial("pcd")[0])
            W0 = m.W.copy()
            qo.persistent_contrastive_divergence(m, X, [range(t * B, (t + 1) * B)], chains, steps=k, learning_rate=lr,
                                                 stream=st, y=Y, label_learning_rate=llr)
            ds.absorb(st.u, "pcd")
            dev.inject_streams(ds.U_h, ds.U_v, ds.U_y)
            dev.pcd_step(t, B, k, lr, llr)
            dv, dh, dy = dev.get_chains()
            ov, oh, oy = chains_to_arrays(chains)
            if bf and L:
                # classifier PCD advances the chains AFTER the update inside the same call (train_pcd.jl:89-92):
                # the device samples with bf16(W_new) while the oracle can only be re-synchronised between
                # calls, so a few draws beyond the guard band may flip; continue from the device's chains
                assert np.mean(dh != oh) < 0.03 and np.mean(dv != ov) < 0.03 and np.mean(dy != oy) < 0.06
                chains[:] = chains_from_arrays(dv, dh, dy)
            else:
                assert np.array_equal(dh, oh) and np.array_equal(dv, ov)
            dev.pull_params()
            if bf:
                assert rel_err(dev.rbm.W - Wprev, m.W - W0) < 2e-2
                Wprev = dev.rbm.W.copy()
                _resync(m, dev, prec)
            else:
                assert rel_err(dev.rbm.W, m.W) < 1e-5 and np.max(np.abs(dev.rbm.a - m.a)) < 1e-6
                assert np.max(np.abs(dev.rbm.b - m.b)) < 1e-6


# ------------------------------------------------------------------------------ quantum-trainer positive phase
@pytest.mark.parametrize("prec", ["fp32", "fp32x3", "bf16"])
@pytest.mark.parametrize("L", [0, 3])
def test_persistent_qubo_minibatch_with_supplied_model_sample(q, prec, L):
    """persistent_qubo! pairs every sample of the mini-batch with ONE (v~, h~) pair returned by the
    annealer; here the pair is supplied by the caller and everything else runs on the GPU."""
    bf = prec == "bf16"
    rng = np.random.default_rng(61 + L)
    V, H, B, steps = 48, 28, 12, 3
    m = make_oracle_model(rng, V, H, L, bf16=bf)
    X, Y = _data(rng, steps * B, V, L, False, bf)
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision=prec) as dev:
        dev.set_data(X, Y)
        Wprev = m.W.copy()
        for t in range(steps):
            v_m = rng.integers(0, 2, V).astype(np.float64)
            h_m = rng.integers(0, 2, H).astype(np.float64)
            y_m = rng.integers(0, 2, L).astype(np.float64) if L else None
            W0 = m.W.copy()
            qo.persistent_qubo_minibatch(m, X, range(t * B, (t + 1) * B), v_m, h_m, learning_rate=0.05, y=Y, y_model=y_m,
                                         label_learning_rate=0.02)
            dev.qsampling_step(t, B, v_m, h_m, 0.05, y_m, 0.02)
            dev.pull_params()
            if bf:
                assert rel_err(dev.rbm.W - Wprev, m.W - W0) < 2e-2
                Wprev = dev.rbm.W.copy()
                _resync(m, dev, prec)
            else:
                assert rel_err(dev.rbm.W, m.W) < 1e-5 and np.max(np.abs(dev.rbm.a - m.a)) < 1e-6
                assert np.max(np.abs(dev.rbm.b - m.b)) < 1e-6
                if L:
                    assert rel_err(dev.rbm.U, m.U) < 1e-5 and np.max(np.abs(dev.rbm.c - m.c)) < 1e-6


def test_device_side_snapshot_and_restore(q):
    rng = np.random.default_rng(71)
    V, H, B = 30, 20, 8
    m = make_oracle_model(rng, V, H, bf16=True)
    X, _ = _data(rng, 2 * B, V, 0, False, True)
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision="bf16", seed=3) as dev:
        dev.set_data(X)
        dev.init_chains(B)
        p0 = dev.conditional_prob_h(X[:4])
        dev.snapshot_params()
        dev.pcd_step(0, B, 1, 0.5)
        dev.pull_params()
        assert np.max(np.abs(dev.rbm.W - m.W)) > 1e-4          # the step really moved the weights
        dev.restore_params()
        dev.pull_params()
        assert np.array_equal(dev.rbm.W, m.W) and np.array_equal(dev.rbm.a, m.a) and np.array_equal(dev.rbm.b, m.b)
        assert np.array_equal(dev.conditional_prob_h(X[:4]), p0)  # shadows were refreshed too


# ------------------------------------------------------------------------------ online (B = 1) CD: cooperative persistent kernel
@pytest.mark.parametrize("prec", ["fp32", "fp32x3", "bf16"])
@pytest.mark.parametrize("V,H,gaussian,k,n", [(3, 2, False, 3, 40), (50, 30, False, 1, 64), (784, 500, False, 2, 24), (24, 16, True, 2, 32)])
def test_online_cd_epoch_matches_reference_algorithm(q, prec, V, H, gaussian, k, n):
    """train!(rbm, x, CD) with the reference's own batch size 1 (src/training/train_cd.jl:2-21): the one-kernel
    epoch must reproduce the oracle's per-sample online algorithm (injected streams, fp32 arithmetic in both
    precision modes) ..."""
    rng = np.random.default_rng(91 + V + k)
    m = make_oracle_model(rng, V, H, 0, gaussian, scale=0.2)
    X, _ = _data(rng, n, V, 0, gaussian, prec == "bf16")
    ds = DenseStreams(rng, k, n, V, H, 0, gaussian)   # one "row" per sample: [k][n][units]
    su, sz = ds.serial("cd")
    st = qo.GuardedStream(su, sz)
    m0 = qo.copy_rbm(m)
    qo.contrastive_divergence(m, X, steps=k, learning_rate=0.02, stream=st, batch_size=1)
    ds.absorb(st.u, "cd")
    with q.DeviceRBM(to_host_rbm(q, m0), max_batch=4, precision=prec) as dev:
        dev.set_data(X)
        dev.inject_streams(ds.U_h, ds.U_v, None, ds.Z_v)
        t = dev.cd_epoch(1, k, 0.02)
        assert all(x >= 0 for x in t) and sum(t) > 0
        dev.pull_params()
        assert rel_err(dev.rbm.W, m.W) < 2e-5 and rel_err(dev.rbm.W - m0.W, m.W - m0.W) < 2e-4
        assert np.max(np.abs(dev.rbm.a - m.a)) < 2e-6 and np.max(np.abs(dev.rbm.b - m.b)) < 2e-6


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("V,H,L,gaussian,k,n", [(3, 2, 1, False, 3, 40), (50, 30, 4, False, 1, 64), (784, 1000, 10, False, 1, 12),
                                                 (24, 16, 3, True, 2, 32)])
def test_online_cd_epoch_classifier_matches_reference_algorithm(q, prec, V, H, L, gaussian, k, n):
    """train!(rbm, x, y, CD) for classifier models (src/training/train_cd.jl:23-43, the classifier default: batch size 1):
    3-argument h_data, chain over (h, v, y) with softmax + independent Bernoulli label draws, h_model from the 2-argument form
    that ignores the sampled labels, W / a / b with learning_rate and U / c with label_learning_rate -- one kernel per epoch
    against the oracle's per-sample algorithm fed the same injected streams."""
    rng = np.random.default_rng(97 + V + k + L)
    m = make_oracle_model(rng, V, H, L, gaussian, scale=0.2)
    X, Y = _data(rng, n, V, L, gaussian, prec == "bf16")
    ds = DenseStreams(rng, k, n, V, H, L, gaussian)
    su, sz = ds.serial("cd")
    st = qo.GuardedStream(su, sz)
    m0 = qo.copy_rbm(m)
    qo.contrastive_divergence(m, X, steps=k, learning_rate=0.02, stream=st, y=Y, label_learning_rate=0.03, batch_size=1)
    ds.absorb(st.u, "cd")
    with q.DeviceRBM(to_host_rbm(q, m0), max_batch=4, precision=prec) as dev:
        dev.set_data(X, Y)
        dev.inject_streams(ds.U_h, ds.U_v, ds.U_y, ds.Z_v)
        dev.reset_counters()
        t = dev.cd_epoch(1, k, 0.02, 0.03)
        assert dev.counters()[0] <= 4, "the epoch must be ONE cooperative launch (+ transposes), not a launch sequence per sample"
        assert all(x >= 0 for x in t) and sum(t) > 0
        dev.pull_params()
        assert rel_err(dev.rbm.W, m.W) < 2e-5 and rel_err(dev.rbm.W - m0.W, m.W - m0.W) < 2e-4
        assert rel_err(dev.rbm.U, m.U) < 2e-5 and rel_err(dev.rbm.U - m0.U, m.U - m0.U) < 2e-4
        assert np.max(np.abs(dev.rbm.a - m.a)) < 2e-6 and np.max(np.abs(dev.rbm.b - m.b)) < 2e-6 and np.max(np.abs(dev.rbm.c - m.c)) < 2e-6


def test_online_cd_classifier_kernel_equals_per_sample_steps(q):
    """... and, with the production Philox stream, the classifier epoch equals a sequence of qb_cd_step(B = 1) calls."""
    rng = np.random.default_rng(94)
    V, H, L, k, n = 64, 40, 5, 2, 48
    m = make_oracle_model(rng, V, H, L, scale=0.2)
    X, Y = _data(rng, n, V, L, False, False)
    out = []
    for online in (1, 0):
        rbm = to_host_rbm(q, m)
        with q.DeviceRBM(rbm, max_batch=4, precision="fp32", seed=23) as dev:
            dev.set_option("online_kernel", online)
            dev.set_data(X, Y)
            dev.cd_epoch(1, k, 0.02, 0.03, timed=False)
            dev.pull_params()
            out.append((rbm.W.copy(), rbm.U.copy(), rbm.a.copy(), rbm.b.copy(), rbm.c.copy()))
    for a, b in zip(*out):
        assert np.max(np.abs(a - b)) < 2e-6
    assert np.max(np.abs(out[0][1] - m.U)) > 1e-3


def test_online_cd_kernel_equals_per_sample_steps(q):
    """... and, with the production Philox stream, equal a sequence of qb_cd_step(B = 1) calls."""
    rng = np.random.default_rng(93)
    V, H, k, n = 64, 40, 2, 48
    m = make_oracle_model(rng, V, H, scale=0.2)
    X, _ = _data(rng, n, V, 0, False, False)
    out = []
    for online in (1, 0):
        rbm = to_host_rbm(q, m)
        with q.DeviceRBM(rbm, max_batch=4, precision="fp32", seed=21) as dev:
            dev.set_option("online_kernel", online)
            dev.set_data(X)
            dev.cd_epoch(1, k, 0.02, timed=False)
            dev.pull_params()
            out.append((rbm.W.copy(), rbm.a.copy(), rbm.b.copy()))
    for a, b in zip(*out):
        assert np.max(np.abs(a - b)) < 2e-6
    assert np.max(np.abs(out[0][0] - m.W)) > 1e-3


# ------------------------------------------------------------------------------ FastPCD
def _fastpcd_oracle(m, X, B, k, lr, flr, seed, cv0, ch0, nb, Y=None, cy0=None, llr=0.0, fllr=0.0, pair=None):
    """Oracle FastPCD epoch fed with the uniforms / normals the device's Philox contract defines."""
    V, H, L = m.n_visible, m.n_hidden, m.n_classifiers
    us, zs = [], []
    for mb in range(nb):
        for s in range(k):
            uh = philox_ref.uniforms(seed, 2 + 2 * (mb * k + s), B, H)
            uv = philox_ref.uniforms(seed, 3 + 2 * (mb * k + s), B, V + L)   # label units ride on the visible draw
            zv = philox_ref.normals(seed, 3 + 2 * (mb * k + s), B, V) if m.gaussian else None
            for i in range(B):
                us += [uh[i], uv[i]]            # serial order per chain: H uniforms, (V normals,) V uniforms, L uniforms
                if m.gaussian:
                    zs.append(zv[i])
    chains = chains_from_arrays(cv0, ch0, cy0)
    stream = qo.Stream(np.concatenate(us), np.concatenate(zs) if zs else None)
    qo.fast_persistent_contrastive_divergence(m, X, qo._set_mini_batches(len(X), B), chains, steps=k, learning_rate=lr,
                                              fast_learning_rate=flr, stream=stream, y=Y, label_learning_rate=llr,
                                              fast_label_learning_rate=fllr, pair_chains=pair)
    return chains_to_arrays(chains)


@pytest.mark.parametrize("V,H,L,gaussian,B,k,nb,pair", [
    (3, 2, 0, False, 2, 1, 6, None), (40, 24, 0, False, 8, 2, 3, None), (784, 500, 0, False, 16, 1, 2, None),
    (40, 24, 5, False, 8, 2, 3, None),      # classifier, reference pairing: every sample with chain 1 (train_fast_pcd.jl:77-119)
    (40, 24, 5, False, 8, 1, 3, True),      # classifier, corrected pairing (option fast_pcd_pair_chains)
    (784, 500, 10, False, 16, 1, 2, None),
    (40, 24, 0, True, 8, 2, 3, None),       # GRBM: the fast chain step thresholds mean + z (gibbs.jl:22-25, rbm.jl:188-189)
    (33, 20, 4, True, 8, 1, 2, None),       # GRBMClassifier
])
def test_fast_pcd_epoch_matches_reference_algorithm(q, V, H, L, gaussian, B, k, nb, pair):
    """fast_persistent_contrastive_divergence! (train_fast_pcd.jl:1-57, classifiers :59-127): per-sample slow update +
    fast-weight decay inside a mini-batch, chains on W + W_fast.  fp32 mode, Philox stream replayed into the oracle."""
    rng = np.random.default_rng(5 + V + L)
    seed, lr, flr, llr, fllr = 1234, 0.05, 0.1, 0.04, 0.08
    m = make_oracle_model(rng, V, H, L, gaussian=gaussian, scale=0.2)
    X, Y = _data(rng, nb * B + 1, V, L, gaussian, False)     # + 1: the trailing remainder is dropped
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=B, precision="fp32", seed=seed) as dev:
        dev.set_data(X, Y)
        if pair:
            dev.set_option("fast_pcd_pair_chains", 1)
        dev.init_chains(B)                                 # draws 0, 1
        cv0, ch0, cy0 = dev.get_chains()
        t = dev.fast_pcd_epoch(B, k, lr, flr, llr, fllr)
        assert t[1] > 0 and t[2] > 0
        dv, dh, dy = dev.get_chains()
        dev.pull_params()
        ov, oh, oy = _fastpcd_oracle(m, X, B, k, lr, flr, seed, cv0, ch0, nb, Y, cy0, llr, fllr, pair)
        assert np.mean(dv != ov) < 2e-3 and np.mean(dh != oh) < 2e-3     # near-ties of the fixed Philox stream only
        if gaussian:
            assert set(np.unique(dv)) <= {0.0, 1.0}        # the reference's fast chains turn a GRBM's visibles binary
        assert rel_err(dev.rbm.W, m.W) < 1e-4 and np.max(np.abs(dev.rbm.b - m.b)) < 1e-4
        assert np.max(np.abs(dev.rbm.a - m.a)) < 1e-4
        if L:
            assert np.mean(dy != oy) < 5e-3
            assert rel_err(dev.rbm.U, m.U) < 1e-4 and np.max(np.abs(dev.rbm.c - m.c)) < 1e-4


def test_cross_entropy_metric_is_reference_literal(q):
    """_evaluate(CrossEntropy) (src/evaluation.jl:22-26) is fed the rounded prediction by `evaluate` (:28-50): every
    term is 0 * -Inf or -Inf, so the reference's metric is NaN for any data; the host mirror reproduces that."""
    rng = np.random.default_rng(3)
    m = make_oracle_model(rng, 12, 6, 3, scale=0.3)
    X, Y = _data(rng, 10, 12, 3, False, False)
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=16, precision="fp32") as dev:
        ce = dev.eval_cross_entropy(X, Y)
        pred = np.stack([qo.classify(m, x) for x in X]).astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            ref = np.sum(Y * np.log(pred) + (1 - Y) * np.log(1 - pred)) / len(X)
        assert np.isnan(ce) and np.isnan(ref)


@pytest.mark.parametrize("L,gaussian", [(0, False), (6, False), (0, True), (6, True)])
def test_fast_pcd_bf16_chains_stay_close_to_fp32(q, L, gaussian):
    rng = np.random.default_rng(6 + L)
    V, H, B, k, nb = 96, 64, 16, 2, 4
    m = make_oracle_model(rng, V, H, L, gaussian=gaussian, scale=0.2, bf16=True)
    X, Y = _data(rng, nb * B, V, L, gaussian, True)
    Ws = {}
    for prec in ("fp32", "bf16"):
        rbm = to_host_rbm(q, m)
        with q.DeviceRBM(rbm, max_batch=B, precision=prec, seed=9) as dev:
            dev.set_data(X, Y)
            dev.set_option("fast_pcd_pair_chains", 1)
            dev.init_chains(B)
            dev.fast_pcd_epoch(B, k, 0.05, 0.1, 0.05, 0.1, timed=False)
            dev.pull_params()
            Ws[prec] = rbm.W.copy()
            cv, ch, cy = dev.get_chains()
            assert set(np.unique(cv)) <= {0.0, 1.0} and set(np.unique(ch)) <= {0.0, 1.0}   # binary even for Gaussian visibles
            # the tensor-core shadows hold the slow weights again after the epoch
            Wr = bf16_round(rbm.W) if prec == "bf16" else rbm.W
            assert np.max(np.abs(dev.conditional_prob_h(X[:2]) - 1 / (1 + np.exp(-(rbm.b + X[:2] @ Wr))))) < 5e-6
    assert rel_err(Ws["bf16"] - m.W, Ws["fp32"] - m.W) < 0.3   # different chain trajectories (bf16-rounded W + W_fast), same learning signal


def test_train_fast_pcd_classifier_runs_like_reference_smoke_test(q, tmp_path):
    """test/classification.jl:40-57 scenario shape: train!(rbm, x, y, FastPCD; ...) on a tiny classifier."""
    rng = np.random.default_rng(2)
    x = [[1, 0, 0, 1]] * 40 + [[0, 1, 1, 0]] * 40
    y = [[1, 0]] * 40 + [[0, 1]] * 40
    rbm = q.RBMClassifier(4, 3, 2, W=rng.standard_normal((4, 3)) * 0.1, U=rng.standard_normal((2, 3)) * 0.1)
    out = q.train(rbm, x, q.FastPCD, y, n_epochs=5, batch_size=4, learning_rate=[0.05] * 5, label_learning_rate=[0.05] * 5,
                  fast_learning_rate=0.05, fast_label_learning_rate=0.05, metrics=[q.Accuracy, q.CrossEntropy], precision="fp32",
                  file_path=str(tmp_path / "m.csv"), verbose=False)
    assert len(out["accuracy"]) == 6 and len(out["cross_entropy"]) == 6
    assert all(np.isnan(v) for v in out["cross_entropy"])        # the reference's metric, literally (see eval_cross_entropy)
    assert (tmp_path / "m.csv").read_text().splitlines()[0] == "accuracy,cross_entropy"
    assert np.all(np.isfinite(rbm.W)) and np.all(np.isfinite(rbm.U))


def test_train_fast_pcd_classifier_reference_scenario(q, tmp_path):
    """test/classification.jl:46-67 `fast_pcd()`: RBMClassifier(3,2,1), x=[1,0,0], y=[1], batch 2, fast rates 0.1, early
    stopping; the reference only checks that it runs and classifies."""
    x = [[1, 0, 0] for _ in range(200)]
    y = [[1] for _ in range(200)]
    rbm = q.RBMClassifier(3, 2, 1, rng=np.random.default_rng(2))
    lr = [0.01 / (j ** 0.8) for j in range(1, 1001)]
    hist = q.train(rbm, x, q.FastPCD, y, n_epochs=5, batch_size=2, learning_rate=lr, label_learning_rate=lr, fast_learning_rate=0.1,
                   fast_label_learning_rate=0.1, early_stopping=True, file_path=str(tmp_path / "f.csv"), precision="fp32", verbose=False)
    assert list(hist) == ["accuracy"] and hist["accuracy"][-1] == 1.0
    with q.DeviceRBM(rbm, max_batch=1, precision="fp32") as dev:
        assert dev.classify([[1, 0, 0]]).tolist() == [[1]]


@pytest.mark.parametrize("prec,L", [("fp32", 0), ("fp32x3", 0), ("bf16", 0), ("fp32", 3)])
def test_step_host_reconstruction_error_value(q, prec, L):
    """qb_pcd_step_host returns sum over rows and visible units of (x - reconstruct(x))^2 (src/rbm.jl:220-224): for binary
    models without labels taken inside the step with the weights BEFORE the update (from the positive-phase p(h|x)),
    otherwise after the step with the updated weights."""
    bf = prec == "bf16"
    rng = np.random.default_rng(77 + L)
    V, H, B, k = 48, 28, 16, 1
    m = make_oracle_model(rng, V, H, L, bf16=bf)
    X, Y = _data(rng, 2 * B, V, L, False, bf)
    rbm = to_host_rbm(q, m)
    with q.DeviceRBM(rbm, max_batch=B, precision=prec, seed=5) as dev:
        dev.init_chains(B)
        for t in range(2):
            sl = slice(t * B, (t + 1) * B)
            dev.pull_params()
            before = qo.Model(rbm.W.copy(), rbm.a.copy(), rbm.b.copy(), None if not L else rbm.U.copy(), None if not L else rbm.c.copy())
            err = dev.step_host("PCD", X[sl], None if Y is None else Y[sl], k, 0.05, 0.02)
            dev.pull_params()
            after = qo.Model(rbm.W.copy(), rbm.a.copy(), rbm.b.copy(), None if not L else rbm.U.copy(), None if not L else rbm.c.copy())
            ref_model = before if L == 0 else after
            if bf:
                ref_model = qo.Model(bf16_round(ref_model.W), ref_model.a, ref_model.b)
            ref = sum(float(np.sum((x - qo.reconstruct(ref_model, x)) ** 2)) for x in X[sl])
            assert abs(err - ref) / ref < (2e-3 if bf else 1e-5), (err, ref)


# ------------------------------------------------------------------------------ single sampling half-steps
@pytest.mark.parametrize("prec", ["fp32", "fp32x3", "bf16"])
@pytest.mark.parametrize("V,H,L,gaussian", [(3, 2, 0, False), (70, 33, 0, False), (40, 24, 5, False), (40, 24, 0, True), (30, 18, 3, True)])
def test_sample_hidden_and_visible_half_steps(q, prec, V, H, L, gaussian):
    """qb_sample_hidden / qb_sample_visible against gibbs_sample_hidden / gibbs_sample_visible / gibbs_sample_label
    (src/gibbs.jl:3-20,29-32,46-49) with the caller's random numbers, then with the Philox stream."""
    bf = prec == "bf16"
    rng = np.random.default_rng(V + H + L)
    n = 13
    m = make_oracle_model(rng, V, H, L, gaussian=gaussian, bf16=bf, scale=0.3)
    X, Y = _data(rng, n, V, L, gaussian, bf)
    Hs = rng.integers(0, 2, (n, H)).astype(np.float64)
    seed = 321
    with q.DeviceRBM(to_host_rbm(q, m), max_batch=8, precision=prec, seed=seed) as dev:     # max_batch < n: chunked
        for use_y in ([False, True] if L else [False]):
            st = qo.GuardedStream(u24(rng, n * H))
            ref = np.stack([qo.gibbs_sample_hidden(m, X[i], st, Y[i] if use_y else None) for i in range(n)])
            out = dev.sample_hidden(X, Y if use_y else None, U=st.u.reshape(n, H))
            assert np.array_equal(out, ref)
        if gaussian:
            Z = f32_round(rng.standard_normal((n, V)))
            st = qo.GuardedStream(u24(rng, n * L) if L else None, Z.reshape(-1))
        else:
            st = qo.GuardedStream(u24(rng, n * (V + L)))
        ref_v, ref_y = [], []
        for i in range(n):
            ref_v.append(np.asarray(qo.gibbs_sample_visible(m, Hs[i], st), dtype=np.float64))
            if L:
                ref_y.append(qo.gibbs_sample_label(m, Hs[i], st))
        if gaussian:
            R, Ry = Z, (st.u.reshape(n, L) if L else None)
        else:
            u = st.u.reshape(n, V + L)
            R, Ry = u[:, :V], (u[:, V:] if L else None)
        res = dev.sample_visible(Hs, R, Ry)
        out_v, out_y = res if L else (res, None)
        if gaussian:
            assert np.max(np.abs(out_v - np.stack(ref_v))) < (5e-2 if bf else 1e-5)     # bf16: the sample is stored in bf16
        else:
            assert np.array_equal(out_v, np.stack(ref_v))
        if L:
            assert np.array_equal(out_y, np.stack(ref_y))
        # Philox: draws 0 (hidden) and 1 (visible + labels) of this handle
        out = dev.sample_hidden(X)
        ph = np.stack([qo.conditional_prob_h(m, x) for x in X])
        uh = philox_ref.uniforms(seed, 0, n, H)
        safe = np.abs(uh - ph) > 1e-4
        assert np.array_equal(out[safe], (uh < ph).astype(np.float64)[safe])
        res = dev.sample_visible(Hs)
        out_v = res[0] if L else res
        if not gaussian:
            pv = np.stack([qo.conditional_prob_v(m, h) for h in Hs])
            uv = philox_ref.uniforms(seed, 1, n, V + L)[:, :V]
            safe = np.abs(uv - pv) > 1e-4
            assert np.array_equal(out_v[safe], (uv < pv).astype(np.float64)[safe])
