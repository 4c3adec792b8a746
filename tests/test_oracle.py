"""CPU tests that pin the oracle's semantics (SURVEY.md section 8(c) items ii, iii, v).

The reference's own tests hold no assertions or vectors, so these are the pins: closed
forms on tiny models, the documented RNG draw order, every quirk listed in section 3.3,
agreement between the two independent restatements (numpy / C), and the equivalence of
the per-sample formulation with the batched-GEMM formulation the device uses.
"""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import qarbom_oracle as qo


def u24(rng, *shape):
    """Uniforms that are exactly representable in fp32 (k / 2^24)."""
    return rng.integers(0, 1 << 24, size=shape).astype(np.float64) / float(1 << 24)


def small_model(rng, V, H, L=0, gaussian=False, scale=0.5):
    W = rng.standard_normal((V, H)) * scale
    U = rng.standard_normal((L, H)) * scale if L else None
    m = qo.make_model(V, H, L, gaussian, W=W, U=U)
    m.a[:] = rng.standard_normal(V) * 0.1
    m.b[:] = rng.standard_normal(H) * 0.1
    if L:
        m.c[:] = rng.standard_normal(L) * 0.1
    return m


# ----------------------------------------------------------------- closed forms
def test_conditionals_closed_form_3x2():
    W = np.array([[0.5, -1.0], [0.25, 2.0], [-0.75, 0.125]])
    m = qo.make_model(3, 2, W=W)
    m.a[:] = [0.1, -0.2, 0.3]
    m.b[:] = [0.05, -0.15]
    v = np.array([1, 0, 1])
    ph = qo.conditional_prob_h(m, v)
    exp_h = [1 / (1 + np.exp(-(0.05 + 0.5 - 0.75))), 1 / (1 + np.exp(-(-0.15 - 1.0 + 0.125)))]
    assert np.allclose(ph, exp_h, rtol=0, atol=1e-15)
    h = np.array([1, 1])
    pv = qo.conditional_prob_v(m, h)
    exp_v = 1 / (1 + np.exp(-(m.a + W.sum(axis=1))))
    assert np.allclose(pv, exp_v, rtol=0, atol=1e-15)


def test_logistic_clamps():
    assert qo.logistic(np.array([-800.0]))[0] == 0.0
    assert qo.logistic(np.array([40.0]))[0] == 1.0
    assert qo.logistic(np.array([0.0]))[0] == 0.5


def test_constructor_draw_order_and_sigma():
    z = np.arange(3 * 2 + 1 * 2, dtype=np.float64)
    m = qo.make_model(3, 2, 1, stream=qo.Stream(normals=z))
    # column-major fill of W (rbm.jl:70), then U (rbm.jl:71); sigma = 1 (no scaling)
    assert np.array_equal(m.W[:, 0], [0, 1, 2]) and np.array_equal(m.W[:, 1], [3, 4, 5])
    assert np.array_equal(m.U[0], [6, 7])
    assert not m.a.any() and not m.b.any() and not m.c.any()
    g = qo.make_model(3, 2, 1, gaussian=True, W=np.zeros((3, 2)), U=np.zeros((1, 2)))
    assert np.all(g.a == 1e-5) and np.all(g.b == 1e-5) and np.all(g.c == 1e-5)  # rbm.jl:94-99


def test_free_energy_matches_enumeration():
    rng = np.random.default_rng(5)
    m = small_model(rng, 4, 3)
    v = np.array([1, 0, 1, 1])
    # F(v) = -log sum_h exp(-E(v,h)),  E from src/qubo.jl:43-47
    tot = 0.0
    for bits in range(8):
        h = np.array([(bits >> j) & 1 for j in range(3)], dtype=np.float64)
        E = -(v @ m.W @ h) - m.a @ v - m.b @ h
        tot += np.exp(-E)
    assert np.isclose(qo.free_energy(m, v), -np.log(tot), rtol=1e-13)


# ----------------------------------------------------------------- draw order / quirks
def test_gibbs_draw_order_cd():
    rng = np.random.default_rng(1)
    m = small_model(rng, 5, 4, 2)
    k = 2
    nu = k * (4 + 5 + 2)
    st = qo.Stream(uniforms=u24(rng, nu))
    v0, y0 = np.array([1, 0, 1, 0, 1]), np.array([0, 1])
    qo._get_v_y_model(m, v0, y0, k, st)
    assert st.upos == nu  # H, V, L uniforms per step, nothing else
    # first H uniforms decide the first hidden sample, in unit order, strict '<'
    st2 = qo.Stream(uniforms=st.u)
    h = qo.gibbs_sample_hidden(m, v0, st2, y0)
    p = qo.conditional_prob_h(m, v0, y0)
    assert np.array_equal(h, (st.u[:4] < p).astype(int))


def test_strict_less_than():
    m = qo.make_model(1, 1, W=np.zeros((1, 1)))
    # p = 0.5 exactly; u == p must give 0 (rand() < p)
    assert qo.gibbs_sample_hidden(m, np.array([1]), qo.Stream(uniforms=[0.5]))[0] == 0
    assert qo.gibbs_sample_hidden(m, np.array([1]), qo.Stream(uniforms=[0.4999999])) [0] == 1


def test_label_sampling_is_independent_bernoulli():
    rng = np.random.default_rng(2)
    m = small_model(rng, 3, 4, 3)
    h = np.array([1, 0, 1, 1])
    p = qo.conditional_prob_y_given_h(m, h)
    assert np.isclose(p.sum(), 1.0)
    assert qo.gibbs_sample_label(m, h, qo.Stream(uniforms=np.zeros(3))).sum() == 3     # all fire
    assert qo.gibbs_sample_label(m, h, qo.Stream(uniforms=np.ones(3) * 0.999999)).sum() == 0  # none


def test_fantasy_init_uniform_floats_and_order():
    m = qo.make_model(3, 2, 1, W=np.zeros((3, 2)), U=np.zeros((1, 2)))
    u = np.arange(2 * 6, dtype=np.float64) / 16
    ch = qo._init_fantasy_data(m, 2, qo.Stream(uniforms=u))
    assert np.array_equal(ch[0].v, u[0:3]) and np.array_equal(ch[0].h, u[3:5]) and np.array_equal(ch[0].y, u[5:6])
    assert np.array_equal(ch[1].v, u[6:9])


def test_chain_h_not_refreshed_and_pcd_uses_binary_h():
    rng = np.random.default_rng(3)
    m = small_model(rng, 6, 5)
    B, k = 4, 2
    st = qo.Stream(uniforms=u24(rng, B * 11 + k * B * 11))
    chains = qo._init_fantasy_data(m, B, st)
    qo._update_fantasy_data(m, chains, k, st)
    for ch in chains:
        assert set(np.unique(ch.h)) <= {0.0, 1.0} and set(np.unique(ch.v)) <= {0.0, 1.0}
    # stored h is the sample that PRODUCED the stored v, not p(h | stored v)
    # (re-running the last half step with the same uniforms reproduces v from h)


def test_set_mini_batches_drops_remainder():
    mb = qo._set_mini_batches(10, 4)
    assert [list(r) for r in mb] == [[0, 1, 2, 3], [4, 5, 6, 7]]
    assert qo._set_mini_batches(8, 4)[-1] == range(4, 8)


def test_mse_is_sum_over_units_over_N():
    m = qo.make_model(3, 2, W=np.zeros((3, 2)))
    x = [[1, 0, 0], [1, 1, 0]]
    # reconstruct = 0.5 everywhere -> each sample contributes 3 * 0.25
    assert np.isclose(qo.evaluate_mse(m, x), 0.75)


def test_cd_classifier_negative_hidden_ignores_y_model():
    rng = np.random.default_rng(4)
    m = small_model(rng, 5, 4, 2)
    x, y = [np.array([1, 0, 1, 1, 0])], [np.array([1, 0])]
    tr = []
    st = qo.Stream(uniforms=u24(rng, 11))
    m0 = qo.copy_rbm(m)
    qo.contrastive_divergence(m, x, steps=1, learning_rate=0.1, stream=st, y=y, label_learning_rate=0.1, trace=tr)
    assert np.allclose(tr[0]["h_model"], qo.conditional_prob_h(m0, tr[0]["v_model"]))  # 2-arg form


def test_classifier_pcd_first_batch_uses_raw_uniform_chains():
    rng = np.random.default_rng(6)
    m = small_model(rng, 4, 3, 2)
    B = 2
    X = rng.integers(0, 2, (B, 4)); Y = np.eye(2)[rng.integers(0, 2, B)]
    st = qo.Stream(uniforms=u24(rng, B * 9 + B * 9))
    chains = qo._init_fantasy_data(m, B, st)
    c0 = [(c.v.copy(), c.h.copy(), c.y.copy()) for c in chains]
    m0 = qo.copy_rbm(m)
    qo.persistent_contrastive_divergence(m, X, [range(0, B)], chains, steps=1, learning_rate=0.1, stream=st, y=Y,
                                         label_learning_rate=0.2)
    dW = sum(np.outer(X[i], qo.conditional_prob_h(m0, X[i], Y[i])) - np.outer(c0[i][0], c0[i][1]) for i in range(B))
    assert np.allclose(m.W, m0.W + 0.1 / B * dW, rtol=0, atol=1e-15)
    dU = sum(np.outer(Y[i], qo.conditional_prob_h(m0, X[i], Y[i])) - np.outer(c0[i][2], c0[i][1]) for i in range(B))
    assert np.allclose(m.U, m0.U + 0.2 / B * dU, rtol=0, atol=1e-15)
    # ... and the chains moved afterwards (binary now)
    assert set(np.unique(chains[0].v)) <= {0.0, 1.0}


def test_diverged_predicates():
    assert not qo._diverged_mse([1.0])
    assert qo._diverged_mse([1.0, 1.5]) and not qo._diverged_mse([1.0, 0.5])
    assert qo._diverged_mse([0.2, 0.9, 0.5])  # improved vs previous but not the minimum
    assert qo._diverged_accuracy([0.5, 0.4]) and not qo._diverged_accuracy([0.5, 0.6])


# ----------------------------------------------------------------- batched == per-sample
def batched_pcd_step(m, X, cv, ch, k, lr, U_h, U_v):
    """The batched-GEMM formulation the device implements."""
    cv, ch, _ = qo.batched_step(m, X, None, "PCD", k, lr, cv=cv, ch=ch, U_h=U_h, U_v=U_v)
    return cv, ch


def test_batched_formulation_equals_per_sample_pcd():
    rng = np.random.default_rng(7)
    V, H, B, k = 24, 16, 8, 3
    m = small_model(rng, V, H, scale=0.3)
    mb = qo.copy_rbm(m)
    X = rng.integers(0, 2, (2 * B, V)).astype(np.float64)
    u_init = u24(rng, B * (V + H))
    U_h, U_v = u24(rng, 2, k, B, H), u24(rng, 2, k, B, V)
    serial = np.concatenate([u_init] + [qo.serial_uniforms_pcd(U_h[t], U_v[t], None) for t in range(2)])
    st = qo.Stream(uniforms=serial)
    chains = qo._init_fantasy_data(m, B, st)
    cv = np.stack([c.v for c in chains]); ch = np.stack([c.h for c in chains])
    qo.persistent_contrastive_divergence(m, X, qo._set_mini_batches(2 * B, B), chains, steps=k, learning_rate=0.05,
                                         stream=st)
    for t in range(2):
        cv, ch = batched_pcd_step(mb, X[t * B:(t + 1) * B], cv, ch, k, 0.05, U_h[t], U_v[t])
    assert st.upos == serial.size
    assert np.array_equal(cv, np.stack([c.v for c in chains]))
    assert np.array_equal(ch, np.stack([c.h for c in chains]))
    assert np.allclose(m.W, mb.W, rtol=0, atol=1e-12) and np.allclose(m.a, mb.a, atol=1e-12)
    assert np.allclose(m.b, mb.b, atol=1e-12)


# ----------------------------------------------------------------- numpy vs C restatement
@pytest.mark.parametrize("L,gaussian", [(0, False), (3, False), (0, True), (2, True)])
def test_numpy_and_c_restatements_agree_pcd(L, gaussian):
    rng = np.random.default_rng(11 + L + 7 * gaussian)
    V, H, B, k, nb = 12, 9, 4, 2, 3
    m = small_model(rng, V, H, L, gaussian, scale=0.4)
    X = rng.standard_normal((nb * B + 1, V)) if gaussian else rng.integers(0, 2, (nb * B + 1, V)).astype(np.float64)
    Y = np.eye(L)[rng.integers(0, L, nb * B + 1)] if L else None
    per_chain = H + (0 if gaussian else V) + L
    u = u24(rng, B * (V + H + L) + nb * k * B * per_chain)
    z = rng.standard_normal(nb * k * B * V)
    st = qo.Stream(uniforms=u, normals=z)
    chains = qo._init_fantasy_data(m, B, st)
    chv = np.stack([c.v for c in chains]).copy(); chh = np.stack([c.h for c in chains]).copy()
    chy = np.stack([c.y for c in chains]).copy() if L else None
    cm = c_oracle.CModel(m.W, m.a, m.b, m.U, m.c, gaussian)
    qo.persistent_contrastive_divergence(m, X, qo._set_mini_batches(len(X), B), chains, steps=k, learning_rate=0.1,
                                         stream=st, y=Y, label_learning_rate=0.05)
    err = c_oracle.pcd_epoch(cm, X, Y, B, k, 0.1, 0.05, chv, chh, chy, u=u[B * (V + H + L):], z=z)
    assert err == 0
    assert np.allclose(cm.W, m.W, rtol=0, atol=1e-13) and np.allclose(cm.a, m.a, atol=1e-13)
    assert np.allclose(cm.b, m.b, atol=1e-13)
    if L:
        assert np.allclose(cm.U, m.U, atol=1e-13) and np.allclose(cm.c, m.c, atol=1e-13)
        assert np.array_equal(chy, np.stack([c.y for c in chains]))
    assert np.allclose(chv, np.stack([c.v for c in chains]), atol=1e-13)
    assert np.array_equal(chh, np.stack([c.h for c in chains]))


@pytest.mark.parametrize("L,gaussian,B", [(0, False, 1), (2, False, 1), (0, True, 1), (0, False, 4), (2, False, 3)])
def test_numpy_and_c_restatements_agree_cd(L, gaussian, B):
    rng = np.random.default_rng(21 + L + B)
    V, H, k, N = 10, 7, 3, 12
    m = small_model(rng, V, H, L, gaussian, scale=0.4)
    X = rng.standard_normal((N, V)) if gaussian else rng.integers(0, 2, (N, V)).astype(np.float64)
    Y = np.eye(L)[rng.integers(0, L, N)] if L else None
    u = u24(rng, N * k * (H + V + L))
    z = rng.standard_normal(N * k * V)
    cm = c_oracle.CModel(m.W, m.a, m.b, m.U, m.c, gaussian)
    qo.contrastive_divergence(m, X, steps=k, learning_rate=0.1, stream=qo.Stream(u, z), y=Y,
                              label_learning_rate=0.05, batch_size=B)
    assert c_oracle.cd_epoch(cm, X, Y, B, k, 0.1, 0.05, u=u, z=z) == 0
    assert np.allclose(cm.W, m.W, rtol=0, atol=1e-13) and np.allclose(cm.b, m.b, atol=1e-13)
    if L:
        assert np.allclose(cm.U, m.U, atol=1e-13)


def test_eval_and_free_energy_agree_numpy_c():
    rng = np.random.default_rng(31)
    m = small_model(rng, 9, 6, 2)
    X = rng.integers(0, 2, (7, 9)).astype(np.float64)
    Y = np.eye(2)[rng.integers(0, 2, 7)]
    cm = c_oracle.CModel(m.W, m.a, m.b, m.U, m.c)
    assert np.isclose(c_oracle.eval_mse(cm, X), qo.evaluate_mse(m, X), rtol=1e-13)
    assert np.isclose(c_oracle.mean_free_energy(cm, X, Y), qo.mean_free_energy(m, X, Y), rtol=1e-13)
    g = small_model(rng, 9, 6, gaussian=True)
    Xg = rng.standard_normal((7, 9)); z = rng.standard_normal(7 * 9)
    cg = c_oracle.CModel(g.W, g.a, g.b, gaussian=True)
    assert np.isclose(c_oracle.eval_mse(cg, Xg, z=z), qo.evaluate_mse(g, Xg, qo.Stream(normals=z)), rtol=1e-13)
    assert np.isclose(c_oracle.mean_free_energy(cg, Xg), qo.mean_free_energy(g, Xg), rtol=1e-13)


# ----------------------------------------------------------------- reference smoke scenario
@pytest.mark.parametrize("method", ["CD", "PCD"])
def test_reference_scenario_learns_100(method):
    """test/generative.jl:4-38 scaled down (200 samples, 15 epochs): after training on
    [1,0,0] the reconstruction prefers unit 1."""
    st = qo.NumpyStream(123)
    m = qo.make_model(3, 2, stream=st)
    x = [[1, 0, 0] for _ in range(200)]
    lr = [0.05 / (j ** 0.8) for j in range(1, 1001)]
    hist = qo.train(m, x, method, n_epochs=15, learning_rate=lr, stream=st, gibbs_steps=3 if method == "CD" else 1,
                    batch_size=2 if method == "PCD" else None, early_stopping=True)
    rec = qo.reconstruct(m, np.array([1, 0, 0]))
    assert rec[0] > 0.5 > max(rec[1], rec[2])
    assert hist["mse"][-1] <= hist["mse"][0]


def test_more_chains_extension_reduces_to_reference_when_equal():
    """The n_chains != batch formula, evaluated at n_chains == batch, is the reference update."""
    rng = np.random.default_rng(41)
    V, H, B, k = 10, 6, 4, 1
    m1 = small_model(rng, V, H)
    m2 = qo.copy_rbm(m1)
    X = rng.integers(0, 2, (B, V)).astype(np.float64)
    u = u24(rng, B * (V + H) + k * B * (V + H))
    s1, s2 = qo.Stream(uniforms=u), qo.Stream(uniforms=u)
    c1, c2 = qo._init_fantasy_data(m1, B, s1), qo._init_fantasy_data(m2, B, s2)
    qo.persistent_contrastive_divergence(m1, X, [range(0, B)], c1, steps=k, learning_rate=0.1, stream=s1)
    qo._pcd_minibatch_more_chains(m2, X, range(0, B), c2, k, 0.1, s2, None, 0.1, 0.0, 0.0, None)
    assert np.allclose(m1.W, m2.W, rtol=0, atol=1e-14) and np.allclose(m1.a, m2.a, atol=1e-14)
    assert np.allclose(m1.b, m2.b, atol=1e-14)


@pytest.mark.parametrize("L", [0, 2])
def test_more_chains_extension_numpy_and_c_agree(L):
    rng = np.random.default_rng(43 + L)
    V, H, B, Nc, k, nb = 9, 6, 4, 10, 2, 2
    m = small_model(rng, V, H, L)
    X = rng.integers(0, 2, (nb * B, V)).astype(np.float64)
    Y = np.eye(L)[rng.integers(0, L, nb * B)] if L else None
    u = u24(rng, Nc * (V + H + L) + nb * k * Nc * (V + H + L))
    st = qo.Stream(uniforms=u)
    chains = qo._init_fantasy_data(m, Nc, st)
    chv = np.stack([c.v for c in chains]).copy(); chh = np.stack([c.h for c in chains]).copy()
    chy = np.stack([c.y for c in chains]).copy() if L else None
    cm = c_oracle.CModel(m.W, m.a, m.b, m.U, m.c)
    qo.persistent_contrastive_divergence(m, X, qo._set_mini_batches(len(X), B), chains, steps=k, learning_rate=0.1,
                                         stream=st, y=Y, label_learning_rate=0.05)
    assert c_oracle.pcd_epoch(cm, X, Y, B, k, 0.1, 0.05, chv, chh, chy, u=u[Nc * (V + H + L):], n_chains=Nc) == 0
    assert np.allclose(cm.W, m.W, rtol=0, atol=1e-13) and np.allclose(cm.b, m.b, atol=1e-13)
    assert np.array_equal(chv, np.stack([c.v for c in chains]))


def test_fast_pcd_classifier_pairs_every_sample_with_chain_one():
    """train_fast_pcd.jl:77-119: `batch_index` is never advanced in the classifier method, so only chain 1 enters the
    update; chains 2..B matter only through the chain advance."""
    rng = np.random.default_rng(0)
    V, H, L, B = 6, 4, 3, 4
    X = rng.integers(0, 2, (B, V)).astype(float)
    Y = np.eye(L)[rng.integers(0, L, B)]

    def run(perturb, pair=None):
        m = qo.make_model(V, H, L, W=np.full((V, H), 0.1), U=np.full((L, H), 0.1))
        r = np.random.default_rng(1)
        chains = [qo.FantasyData(r.random(V), r.random(H), r.random(L)) for _ in range(B)]
        if perturb:
            for c in chains[1:]:
                c.v[:] = 1 - c.v
                c.h[:] = 1 - c.h
                c.y[:] = 1 - c.y
        qo.fast_persistent_contrastive_divergence(m, X, [range(B)], chains, steps=1, learning_rate=0.1, fast_learning_rate=0.1,
                                                  stream=qo.NumpyStream(3), y=Y, label_learning_rate=0.1,
                                                  fast_label_learning_rate=0.1, pair_chains=pair)
        return m
    a, b = run(False), run(True)
    assert np.array_equal(a.W, b.W) and np.array_equal(a.U, b.U) and np.array_equal(a.c, b.c)
    c, d = run(False, True), run(True, True)
    assert not np.array_equal(c.W, d.W)       # the corrected pairing does depend on chains 2..B


def test_fast_pcd_gaussian_chains_turn_binary():
    """src/gibbs.jl:22-25 has one fast `gibbs_sample_visible` for every model: it thresholds rand() < mu + randn()."""
    rng = np.random.default_rng(0)
    V, H, B = 5, 3, 4
    m = qo.make_model(V, H, 0, True, W=rng.standard_normal((V, H)) * 0.1)
    X = rng.standard_normal((B, V))
    chains = [qo.FantasyData(rng.random(V), rng.random(H)) for _ in range(B)]
    qo.fast_persistent_contrastive_divergence(m, X, [range(B)], chains, steps=1, learning_rate=0.1, fast_learning_rate=0.1,
                                              stream=qo.NumpyStream(3))
    assert all(set(np.unique(c.v)) <= {0.0, 1.0} for c in chains)


@pytest.mark.parametrize("algo", ["PCD", "CD"])
@pytest.mark.parametrize("L,gaussian", [(0, False), (3, False), (0, True), (2, True)])
def test_batched_step_equals_per_sample_restatement(algo, L, gaussian):
    """`batched_step` (BLAS-3; also bench.py's best-case CPU figure) against the per-sample restatement on the same streams."""
    rng = np.random.default_rng(5 + L + 3 * gaussian)
    V, H, B, k = 10, 7, 6, 2
    m = small_model(rng, V, H, L, gaussian, scale=0.4)
    mb = qo.copy_rbm(m)
    X = rng.standard_normal((B, V)) if gaussian else rng.integers(0, 2, (B, V)).astype(np.float64)
    Y = np.eye(L)[rng.integers(0, L, B)] if L else None
    U_h = u24(rng, k, B, H)
    U_v = None if gaussian else u24(rng, k, B, V)
    Z_v = rng.standard_normal((k, B, V)) if gaussian else None
    U_y = u24(rng, k, B, L) if L else None
    cv, ch = rng.random((B, V)), rng.random((B, H))
    cy = rng.random((B, L)) if L else None
    if algo == "PCD":
        chains = [qo.FantasyData(cv[i].copy(), ch[i].copy(), None if cy is None else cy[i].copy()) for i in range(B)]
        st = qo.Stream(qo.serial_uniforms_pcd(U_h, U_v, U_y), None if Z_v is None else qo.serial_normals(Z_v, "pcd"))
        qo.persistent_contrastive_divergence(m, X, [range(B)], chains, steps=k, learning_rate=0.05, stream=st, y=Y,
                                             label_learning_rate=0.03)
        cv2, ch2, cy2 = qo.batched_step(mb, X, Y, "PCD", k, 0.05, 0.03, cv=cv, ch=ch, cy=cy, U_h=U_h, U_v=U_v, U_y=U_y, Z_v=Z_v)
        assert np.allclose(cv2, np.stack([c.v for c in chains]), atol=1e-12) and np.array_equal(ch2, np.stack([c.h for c in chains]))
    else:
        st = qo.Stream(qo.serial_uniforms_cd(U_h, U_v, U_y), None if Z_v is None else qo.serial_normals(Z_v, "cd"))
        qo.contrastive_divergence(m, X, steps=k, learning_rate=0.05, stream=st, y=Y, label_learning_rate=0.03, batch_size=B)
        qo.batched_step(mb, X, Y, "CD", k, 0.05, 0.03, U_h=U_h, U_v=U_v, U_y=U_y, Z_v=Z_v)
    assert np.allclose(m.W, mb.W, atol=1e-12) and np.allclose(m.a, mb.a, atol=1e-12) and np.allclose(m.b, mb.b, atol=1e-12)
    if L:
        assert np.allclose(m.U, mb.U, atol=1e-12) and np.allclose(m.c, mb.c, atol=1e-12)


@pytest.mark.parametrize("L,gaussian,pair", [(0, False, None), (3, False, None), (3, False, True), (0, True, None), (2, True, None)])
def test_numpy_and_c_restatements_agree_fast_pcd(L, gaussian, pair):
    """Two independent restatements of fast_persistent_contrastive_divergence! (train_fast_pcd.jl:1-127) on the same streams."""
    rng = np.random.default_rng(41 + L + 5 * gaussian)
    V, H, B, k, nb = 11, 7, 4, 2, 3
    m = small_model(rng, V, H, L, gaussian, scale=0.4)
    X = rng.standard_normal((nb * B + 1, V)) if gaussian else rng.integers(0, 2, (nb * B + 1, V)).astype(np.float64)
    Y = np.eye(L)[rng.integers(0, L, nb * B + 1)] if L else None
    u = u24(rng, nb * k * B * (H + V + L))
    z = rng.standard_normal(nb * k * B * V) if gaussian else None
    chains = [qo.FantasyData(rng.random(V), rng.random(H), rng.random(L) if L else None) for _ in range(B)]
    chv = np.stack([c.v for c in chains]).copy(); chh = np.stack([c.h for c in chains]).copy()
    chy = np.stack([c.y for c in chains]).copy() if L else None
    cm = c_oracle.CModel(m.W, m.a, m.b, m.U, m.c, gaussian)
    qo.fast_persistent_contrastive_divergence(m, X, qo._set_mini_batches(len(X), B), chains, steps=k, learning_rate=0.1,
                                              fast_learning_rate=0.2, stream=qo.Stream(u, z), y=Y, label_learning_rate=0.05,
                                              fast_label_learning_rate=0.15, pair_chains=pair)
    assert c_oracle.fast_pcd_epoch(cm, X, Y, B, k, 0.1, 0.2, 0.05, 0.15, chv, chh, chy, u=u, z=z, pair_chains=pair) == 0
    assert np.allclose(cm.W, m.W, rtol=0, atol=1e-13) and np.allclose(cm.a, m.a, atol=1e-13) and np.allclose(cm.b, m.b, atol=1e-13)
    if L:
        assert np.allclose(cm.U, m.U, atol=1e-13) and np.allclose(cm.c, m.c, atol=1e-13)
        assert np.array_equal(chy, np.stack([c.y for c in chains]))
    assert np.array_equal(chv, np.stack([c.v for c in chains])) and np.array_equal(chh, np.stack([c.h for c in chains]))
