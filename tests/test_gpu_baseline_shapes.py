"""Parity of the BENCHMARKED kernels with the Float64 oracle at BASELINE.json's shapes (run with -m gpu).

The small-shape parity tests inject streams and therefore exercise the generic epilogue on tiny tiles.  What
bench.py times is something else: the cta_group::2 256 x 256 kernels with the compile-time epilogues and the
Philox stream.  These tests pin exactly those kernels (asserted through qb_get_gemm_histogram) to the oracle:

  * cfg5  Binary RBM 4096 x 4096, PCD, batch 8192 (k = 1 step by step, and k = 10 in one call),
  * cfg3  Gaussian-Bernoulli RBM 1024 x 2048, batch 1024 (half-steps, and a CD-5 step),
  * cfg4  RBMClassifier 784 + 10 x 1000, batch 512 (CD-1 step, and the label half-step).

Rows of a mini-batch are independent given W (src/fantasy_data.jl:12-19, src/training/train_pcd.jl:22-39), so
the oracle checks a SUBSET of chain rows of the full-size device step; every unit of those rows is compared.
The oracle is fed the uniforms / normals of the stream contract (tests/philox_ref.py) and the bf16-rounded
weights the tensor cores see.  Sampled states must agree bit for bit except where a uniform lies within a guard
band of its probability; the band is 4x the MEASURED probability error of the device (itself asserted < 5e-6),
not a free parameter.  Parameter increments: 2e-2 relative in bf16 mode (probabilities enter the gradient GEMM
rounded to bf16), 5e-2 with Gaussian visibles (real-valued states are stored in bf16).
"""
import numpy as np
import pytest

from oracle import qarbom_oracle as qo
from tests import philox_ref
from tests.helpers import bf16_round, f32_round, rel_err

pytestmark = pytest.mark.gpu

SEED = 0x5EED0B200


@pytest.fixture(scope="module")
def q():
    import qarbom_b200
    qarbom_b200.build()
    return qarbom_b200


def _p_hidden(m, v, y=None):
    """conditional_prob_h for a block of rows (src/rbm.jl:165,170-171), the batched form of qo.conditional_prob_h."""
    pre = m.b + v @ m.W
    if y is not None:
        pre = pre + y @ m.U
    return qo.logistic(pre)


def _mean_visible(m, h):
    """a + W h for a block of rows (src/rbm.jl:182,187)."""
    return m.a + h @ m.W.T


def _check_bernoulli(name, dev_state, p, u, guard):
    """States must equal 1[u < p] except inside the guard band; returns the number of in-band disagreements."""
    exp = (u < p).astype(np.float64)
    mism = exp != dev_state
    n = int(mism.sum())
    if n:
        worst = float(np.max(np.abs(u - p)[mism]))
        assert worst < guard, f"{name}: {n} states differ from the oracle, farthest |u - p| = {worst:.3e} >= guard {guard:.1e}"
    # the expected number of draws inside the band is 2 * guard * draws; disagreements cannot exceed that by much
    assert n <= 10 + 8 * guard * exp.size, f"{name}: {n} in-band disagreements out of {exp.size} draws"
    return n


def _sync_from_device(m, dev):
    dev.pull_params()
    r = dev.rbm
    m.W[...] = bf16_round(r.W)
    m.a[...] = r.a
    m.b[...] = r.b
    if m.is_classifier:
        m.U[...] = bf16_round(r.U)
        m.c[...] = r.c


def _probability_error(dev, m, rows_v, rows_y=None):
    got = dev.conditional_prob_h(rows_v, rows_y)
    return float(np.max(np.abs(got - _p_hidden(m, rows_v, rows_y))))


# ------------------------------------------------------------------------------------------------ cfg5
V5 = H5 = 4096
B5 = 8192
# contiguous row blocks: first / last rows of the batch, a block straddling a 256-row tile boundary, a block straddling the
# boundary between two CTA-pair waves, and two arbitrary ones -- every one of the 4096 + 4096 units of these rows is compared
BLOCKS5 = [(0, 16), (248, 16), (3000, 16), (4090, 12), (6661, 16), (8176, 16)]


@pytest.fixture(scope="module")
def cfg5():
    rng = np.random.default_rng(55)
    m = qo.make_model(V5, H5, W=bf16_round(0.01 * rng.standard_normal((V5, H5))))   # bench.py's W0 law, bf16-exact
    m.a[:] = f32_round(0.05 * rng.standard_normal(V5))
    m.b[:] = f32_round(0.05 * rng.standard_normal(H5))
    X = (rng.random((B5, V5), dtype=np.float32) < 0.5).astype(np.uint8)
    cv0 = rng.integers(0, 2, (B5, V5)).astype(np.float64)    # binary initial chains: exact in bf16
    ch0 = rng.integers(0, 2, (B5, H5)).astype(np.float64)
    rows = np.concatenate([np.arange(a, a + n) for a, n in BLOCKS5])
    return m, X, cv0, ch0, rows


def _rows_of(dev, blocks):
    vs, hs = [], []
    for a, n in blocks:
        v, h, _ = dev.get_chain_rows(a, n)
        vs.append(v); hs.append(h)
    return np.concatenate(vs), np.concatenate(hs)


def _host_rbm(q, m):
    r = q.RBM(m.n_visible, m.n_hidden, W=m.W) if not m.gaussian else q.GRBM(m.n_visible, m.n_hidden, W=m.W)
    r.a[:] = m.a
    r.b[:] = m.b
    return r


def _assert_kernels(hist, fused, width, per_launch):
    """Which kernels produced the step: the persistent multi-half-step kernel (+ the gradient GEMM), or -- option
    fused_chain = 0 -- one launch per contraction with the lean compile-time epilogues."""
    if fused:
        assert hist.get(("2cta", width, "chain")) == 1 and hist.get(("2cta", 256, "grad")) == 1 and sum(hist.values()) == 2, hist
    else:
        for key, n in per_launch.items():
            assert hist.get(("2cta", width, key)) == n, hist
        assert sum(hist.values()) == sum(per_launch.values()), hist


@pytest.mark.parametrize("fused", [2, 0])
def test_cfg5_pcd1_half_steps_and_update_match_oracle(q, cfg5, fused):
    m0, X, cv0, ch0, rows = cfg5
    m = qo.copy_rbm(m0)
    lr = 0.01
    with q.DeviceRBM(_host_rbm(q, m), max_batch=B5, precision="bf16", seed=SEED) as dev:
        dev.set_option("fused_chain", fused)
        dev.set_data(X)
        dev.init_chains(B5, cv0, ch0)
        eps = _probability_error(dev, m, X[rows].astype(np.float64))
        assert eps < 5e-6, f"device p(h|v) error {eps:.2e}"
        guard = max(4 * eps, 2e-6)
        v_prev, draw, inband = cv0[rows], 0, 0
        for t in range(3):
            dev.reset_counters()
            dev.pcd_step(0, B5, 1, lr)
            hist = dev.gemm_histogram()
            # the benchmarked kernels: pair kernel, 256-wide tiles, lean sampling / probability / gradient epilogues
            _assert_kernels(hist, fused, 256, {"fast_sample": 2, "fast_prob": 1, "grad": 1})
            dv, dh = _rows_of(dev, BLOCKS5)
            # gibbs_sample_hidden (src/gibbs.jl:3-6) from the device's own previous visible states ...
            inband += _check_bernoulli(f"step {t} hidden", dh, _p_hidden(m, v_prev), philox_ref.uniforms(SEED, draw, rows, H5), guard)
            # ... and gibbs_sample_visible (src/gibbs.jl:13-16) from the device's own hidden states
            inband += _check_bernoulli(f"step {t} visible", dv, qo.logistic(_mean_visible(m, dh)),
                                       philox_ref.uniforms(SEED, draw + 1, rows, V5), guard)
            draw += 2
            v_prev = dv
            if t == 0:
                # parameter increments of a subset of hidden units (all visible units), all 8192 rows:
                # delta_W = lr/B (x h_data' - v_chain h_chain'), src/training/train_pcd.jl:30-39
                J = np.r_[0:8, 120:136, 1000:1016, 4088:4096, np.random.default_rng(1).choice(H5, 16, replace=False)]
                Xf = X.astype(np.float64)
                hd = qo.logistic(m.b[J] + Xf @ m.W[:, J])
                cv, ch, _ = dev.get_chains()
                assert np.array_equal(cv[rows], dv) and np.array_equal(ch[rows], dh)
                dW = (lr / B5) * (Xf.T @ hd - cv.T @ ch[:, J])
                da = (lr / B5) * (Xf.sum(0) - cv.sum(0))
                db = (lr / B5) * (hd.sum(0) - ch[:, J].sum(0))
                W_before, a_before, b_before = m.W.copy(), m.a.copy(), m.b.copy()
                _sync_from_device(m, dev)
                r = dev.rbm
                assert rel_err(r.W[:, J] - W_before[:, J], dW) < 2e-2
                assert rel_err(r.a - a_before, da) < 1e-3 and rel_err(r.b[J] - b_before[J], db) < 1e-3
                del Xf, cv, ch
            else:
                _sync_from_device(m, dev)
    print(f"cfg5 PCD-1: 3 steps x {rows.size} rows x {V5 + H5} units, guard {guard:.1e}, {inband} in-band disagreements")


@pytest.mark.parametrize("fused", [2, 0])
def test_cfg5_pcd10_single_call_matches_oracle_on_row_subset(q, cfg5, fused):
    m, X, cv0, ch0, rows = cfg5
    k = 10
    with q.DeviceRBM(_host_rbm(q, m), max_batch=B5, precision="bf16", seed=SEED) as dev:
        dev.set_option("fused_chain", fused)
        dev.set_data(X)
        dev.init_chains(B5, cv0, ch0)
        eps = _probability_error(dev, m, X[rows].astype(np.float64))
        assert eps < 5e-6
        guard = max(4 * eps, 2e-6)
        dev.reset_counters()
        dev.pcd_step(0, B5, k, 0.01)      # the chains advance BEFORE the update (train_pcd.jl:14-16): W is m.W throughout
        _assert_kernels(dev.gemm_histogram(), fused, 256, {"fast_sample": 2 * k, "fast_prob": 1, "grad": 1})
        dv, dh = _rows_of(dev, BLOCKS5)
    # _update_fantasy_data! (src/fantasy_data.jl:12-19) for the row subset, remembering how close every draw came to its p
    v = cv0[rows]
    margin = np.full(rows.size, np.inf)
    for s in range(k):
        ph = _p_hidden(m, v)
        uh = philox_ref.uniforms(SEED, 2 * s, rows, H5)
        h = (uh < ph).astype(np.float64)
        pv = qo.logistic(_mean_visible(m, h))
        uv = philox_ref.uniforms(SEED, 2 * s + 1, rows, V5)
        v = (uv < pv).astype(np.float64)
        margin = np.minimum(margin, np.minimum(np.abs(uh - ph).min(1), np.abs(uv - pv).min(1)))
    clean = margin > guard          # rows whose whole 10-step trajectory is decided outside the guard band
    assert clean.mean() >= 0.25, f"only {clean.sum()} of {rows.size} rows stay outside the guard band: inconclusive"
    assert np.array_equal(dh[clean], h[clean]) and np.array_equal(dv[clean], v[clean]), \
        "chain states after 10 Gibbs steps differ from the oracle on rows with no near-tie"
    # rows with a near-tie somewhere: still close (a flipped unit perturbs later probabilities only slightly)
    if (~clean).any():
        assert np.mean(dv[~clean] != v[~clean]) < 0.02
    # the last visible half-step, from the device's own hidden states: every row, every unit
    _check_bernoulli("last visible half-step", dv, qo.logistic(_mean_visible(m, dh)),
                     philox_ref.uniforms(SEED, 2 * k - 1, rows, V5), guard)
    print(f"cfg5 PCD-10: {int(clean.sum())}/{rows.size} rows bit-exact over all 10 steps (guard {guard:.1e})")


# ------------------------------------------------------------------------------------------------ cfg3
V3, H3, B3, K3 = 1024, 2048, 1024, 5


@pytest.fixture(scope="module")
def cfg3():
    rng = np.random.default_rng(33)
    m = qo.make_model(V3, H3, gaussian=True, W=bf16_round(0.01 * rng.standard_normal((V3, H3))))
    m.a[:] = f32_round(0.05 * rng.standard_normal(V3))
    m.b[:] = f32_round(0.05 * rng.standard_normal(H3))
    X = bf16_round(rng.standard_normal((2 * B3, V3)))          # zero-mean / unit-variance data, bf16-exact
    return m, X


def test_cfg3_gaussian_half_steps_match_oracle(q, cfg3):
    """gibbs_sample_hidden + the Gaussian conditional_prob_v (= a sample, src/rbm.jl:187) at the cfg3 shape, bf16 mode,
    Philox normals, every row and unit, each half-step checked from the device's own input states."""
    m0, X = cfg3
    m = qo.copy_rbm(m0)
    rng = np.random.default_rng(3)
    cv0 = bf16_round(rng.standard_normal((B3, V3)))
    ch0 = rng.integers(0, 2, (B3, H3)).astype(np.float64)
    with q.DeviceRBM(_host_rbm(q, m), max_batch=B3, precision="bf16", seed=SEED) as dev:
        dev.set_data(X)
        dev.init_chains(B3, cv0, ch0)
        eps = _probability_error(dev, m, X[:64])
        assert eps < 5e-6
        guard = max(4 * eps, 2e-6)
        v_prev, draw = cv0, 0
        for t in range(3):
            dev.pcd_step(t % 2, B3, 1, 0.002)
            dv, dh, _ = dev.get_chains()
            _check_bernoulli(f"step {t} hidden", dh, _p_hidden(m, v_prev), philox_ref.uniforms(SEED, draw, B3, H3), guard)
            real = _mean_visible(m, dh) + philox_ref.normals(SEED, draw + 1, B3, V3)
            # the device rounds mean + z to bf16 when it stores the state: half a bf16 ulp (at most 2^-8 relative) plus the fp32
            # evaluation error of the mean and of Box-Muller
            err = np.abs(dv - real)
            assert np.all(err <= np.abs(real) * 2.0 ** -8 + 2e-5), f"step {t}: visible sample off by {err.max():.3e}"
            assert np.mean(dv == bf16_round(real)) > 0.99     # and it is the SAME rounding almost everywhere
            v_prev, draw = dv, draw + 2
            _sync_from_device(m, dev)


@pytest.mark.parametrize("fused", [2, 0])
def test_cfg3_cd5_step_matches_oracle(q, cfg3, fused):
    """Two CD-5 mini-batch steps at the cfg3 shape against the oracle's batched formulation (tested equal to the
    per-sample restatement of src/training/train_cd.jl:2-21) fed the same Philox numbers, with the oracle storing
    Gaussian states in bf16 like the device."""
    m0, X = cfg3
    m = qo.copy_rbm(m0)
    lr = 0.002
    with q.DeviceRBM(_host_rbm(q, m), max_batch=B3, precision="bf16", seed=SEED) as dev:
        dev.set_option("fused_chain", fused)
        dev.set_data(X)
        draw = 0
        for t in range(2):
            U_h = np.stack([philox_ref.uniforms(SEED, draw + 2 * s, B3, H3) for s in range(K3)])
            Z_v = np.stack([philox_ref.normals(SEED, draw + 2 * s + 1, B3, V3) for s in range(K3)])
            draw += 2 * K3
            W0, a0, b0 = m.W.copy(), m.a.copy(), m.b.copy()
            qo.batched_step(m, X[t * B3:(t + 1) * B3], None, "CD", K3, lr, U_h=U_h, Z_v=Z_v, v_store=bf16_round)
            dev.reset_counters()
            dev.cd_step(t, B3, K3, lr)
            hist = dev.gemm_histogram()
            if fused:   # 2k+1 forward contractions in one persistent launch + the gradient GEMM
                assert sum(n for key, n in hist.items() if key[2] == "chain") == 1 and sum(hist.values()) == 2, hist
            else:       # the Gaussian visible half-steps run the compile-time Box-Muller epilogue, not the generic one
                assert sum(hist.values()) == 2 * K3 + 2 and not any(key[2] == "generic" or key[0] == "simt" for key in hist), hist
                assert sum(n for key, n in hist.items() if key[2] == "fast_gauss") == K3, hist
            Wd0 = dev.rbm.W.copy()      # the device's fp32 master before the step (the oracle restarts from its bf16 rounding)
            dev.pull_params()
            r = dev.rbm
            assert rel_err(r.W - Wd0, m.W - W0) < 5e-2, f"step {t}: dW rel err {rel_err(r.W - Wd0, m.W - W0):.3e}"
            assert rel_err(r.a - a0, m.a - a0) < 5e-2 and rel_err(r.b - b0, m.b - b0) < 5e-2
            _sync_from_device(m, dev)


# ------------------------------------------------------------------------------------------------ cfg4
V4, H4, L4, B4 = 784, 1000, 10, 512


@pytest.fixture(scope="module")
def cfg4():
    rng = np.random.default_rng(44)
    m = qo.make_model(V4, H4, L4, W=bf16_round(0.01 * rng.standard_normal((V4, H4))),
                      U=bf16_round(0.01 * rng.standard_normal((L4, H4))))
    m.a[:] = f32_round(0.05 * rng.standard_normal(V4))
    m.b[:] = f32_round(0.05 * rng.standard_normal(H4))
    m.c[:] = f32_round(0.05 * rng.standard_normal(L4))
    n = 2 * B4
    X = (rng.random((n, V4)) < 0.1307).astype(np.float64)
    Y = np.eye(L4)[np.arange(n) % L4]
    return m, X, Y


def _host_clf(q, m):
    r = q.RBMClassifier(m.n_visible, m.n_hidden, m.n_classifiers, W=m.W, U=m.U)
    r.a[:] = m.a; r.b[:] = m.b; r.c[:] = m.c
    return r


@pytest.mark.parametrize("fused", [2, 0])
def test_cfg4_classifier_cd1_step_matches_oracle(q, cfg4, fused):
    """Two CD-1 steps of the RBMClassifier at the cfg4 shape (src/training/train_cd.jl:23-43 with the mini-batch extension):
    3-arg positive phase, label units sampled as independent Bernoullis of the softmax, 2-arg negative hidden."""
    m0, X, Y = cfg4
    m = qo.copy_rbm(m0)
    lr, llr = 0.01, 0.02
    with q.DeviceRBM(_host_clf(q, m), max_batch=B4, precision="bf16", seed=SEED) as dev:
        dev.set_option("fused_chain", fused)
        dev.set_data(X, Y)
        draw = 0
        for t in range(2):
            U_h = philox_ref.uniforms(SEED, draw, B4, H4)[None]
            U_vy = philox_ref.uniforms(SEED, draw + 1, B4, V4 + L4)      # label units are units V .. V+L-1 of the visible draw
            draw += 2
            W0, U0, a0, b0, c0 = m.W.copy(), m.U.copy(), m.a.copy(), m.b.copy(), m.c.copy()
            sl = slice(t * B4, (t + 1) * B4)
            qo.batched_step(m, X[sl], Y[sl], "CD", 1, lr, llr, U_h=U_h, U_v=U_vy[None, :, :V4], U_y=U_vy[None, :, V4:])
            dev.reset_counters()
            dev.cd_step(t, B4, 1, lr, llr)
            hist = dev.gemm_histogram()      # label units are finished inside the visible GEMM's epilogue, no side kernel
            assert not any(key[2] == "generic" for key in hist), hist
            if fused:
                assert sum(n for key, n in hist.items() if key[2] == "chain") == 1 and sum(hist.values()) == 2, hist
                assert t == 0 or dev.counters()[0] <= 3, dev.counters()   # chain + gradient + update
            else:
                assert sum(n for key, n in hist.items() if key[2] == "fast_label") == 1 and sum(hist.values()) == 4, hist
                assert t == 0 or dev.counters()[0] <= 5, dev.counters()   # (the first step also builds the transposed data copies)
            Wd0, Ud0 = dev.rbm.W.copy(), dev.rbm.U.copy()   # the device's fp32 masters before the step
            dev.pull_params()
            r = dev.rbm
            assert rel_err(r.W - Wd0, m.W - W0) < 2e-2, f"step {t}: dW rel err {rel_err(r.W - Wd0, m.W - W0):.3e}"
            assert rel_err(r.U - Ud0, m.U - U0) < 2e-2, f"step {t}: dU rel err {rel_err(r.U - Ud0, m.U - U0):.3e}"
            assert rel_err(r.a - a0, m.a - a0) < 2e-2 and rel_err(r.b - b0, m.b - b0) < 2e-2
            # label biases are sums of exact 0/1 differences: at most one near-tie flip of difference
            assert np.max(np.abs((r.c - c0) - (m.c - c0))) <= 1.5 * llr / B4
            _sync_from_device(m, dev)


def test_cfg4_classifier_chain_half_steps_match_oracle(q, cfg4):
    """gibbs_sample_hidden(v, y), gibbs_sample_visible and gibbs_sample_label (src/gibbs.jl:29-32,13-16,46-49) at the cfg4
    shape through a classifier PCD-1 step, whose chains advance AFTER the update (src/training/train_pcd.jl:89-92)."""
    m0, X, Y = cfg4
    m = qo.copy_rbm(m0)
    rng = np.random.default_rng(4)
    cv0 = rng.integers(0, 2, (B4, V4)).astype(np.float64)
    ch0 = rng.integers(0, 2, (B4, H4)).astype(np.float64)
    cy0 = rng.integers(0, 2, (B4, L4)).astype(np.float64)
    with q.DeviceRBM(_host_clf(q, m), max_batch=B4, precision="bf16", seed=SEED) as dev:
        dev.set_data(X, Y)
        dev.init_chains(B4, cv0, ch0, cy0)
        v_prev, y_prev, draw = cv0, cy0, 0
        for t in range(2):
            dev.pcd_step(t, B4, 1, 0.01, 0.02)
            _sync_from_device(m, dev)                  # the chains were advanced with the UPDATED parameters
            if t == 0:
                eps = _probability_error(dev, m, X[:64], Y[:64])
                assert eps < 5e-6
                guard = max(4 * eps, 2e-6)
            dv, dh, dy = dev.get_chains()
            _check_bernoulli(f"step {t} hidden", dh, _p_hidden(m, v_prev, y_prev), philox_ref.uniforms(SEED, draw, B4, H4), guard)
            u = philox_ref.uniforms(SEED, draw + 1, B4, V4 + L4)
            _check_bernoulli(f"step {t} visible", dv, qo.logistic(_mean_visible(m, dh)), u[:, :V4], guard)
            py = np.stack([qo.conditional_prob_y_given_h(m, hrow) for hrow in dh])
            _check_bernoulli(f"step {t} label", dy, py, u[:, V4:], 16 * guard)   # softmax in fp32: exp / log, looser band
            v_prev, y_prev, draw = dv, dy, draw + 2
