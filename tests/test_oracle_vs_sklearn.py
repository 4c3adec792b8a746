"""Independent cross-check of the oracle's shared RBM arithmetic against scikit-learn's BernoulliRBM (a third-party
implementation written from the same equations): p(h|v), p(v|h), the free energy, and a whole k-step Gibbs chain
replayed bit-for-bit from the same uniform stream.  This does NOT pin the reference's training semantics
(scikit-learn's PCD variant differs from QARBoM.jl's) -- it guards the building blocks the
restatement and the CUDA path share."""
import numpy as np
import pytest

from oracle import qarbom_oracle as qo

sk = pytest.importorskip("sklearn.neural_network")


def _pair(rng, V, H):
    m = qo.make_model(V, H, W=rng.standard_normal((V, H)) * 0.4)
    m.a[:] = rng.standard_normal(V) * 0.2
    m.b[:] = rng.standard_normal(H) * 0.2
    r = sk.BernoulliRBM(n_components=H)
    r.components_ = np.ascontiguousarray(m.W.T)      # [H][V] == the C view of Julia's column-major W
    r.intercept_visible_ = m.a.copy()
    r.intercept_hidden_ = m.b.copy()
    r.n_features_in_ = V
    return m, r


@pytest.mark.parametrize("V,H", [(3, 2), (16, 8), (70, 33)])
def test_conditionals_and_free_energy(V, H):
    rng = np.random.default_rng(V)
    m, r = _pair(rng, V, H)
    X = rng.integers(0, 2, (11, V)).astype(np.float64)
    ph = np.stack([qo.conditional_prob_h(m, x) for x in X])
    assert np.max(np.abs(ph - r._mean_hiddens(X))) < 1e-15
    Hs = rng.integers(0, 2, (11, H)).astype(np.float64)
    pv = np.stack([qo.conditional_prob_v(m, h) for h in Hs])
    from scipy.special import expit
    assert np.max(np.abs(pv - expit(Hs @ r.components_ + r.intercept_visible_))) < 1e-15
    F = np.array([qo.free_energy(m, x) for x in X])
    assert np.max(np.abs(F - r._free_energy(X))) < 1e-12
    assert abs(qo.mean_free_energy(m, list(X)) - r._free_energy(X).mean()) < 1e-12


@pytest.mark.parametrize("k", [1, 4])
def test_gibbs_chain_replays_bit_for_bit(k):
    """scikit-learn draws rng.uniform(size=(n, H)) then (n, V) per step and thresholds with strict `<`, like
    src/gibbs.jl:3-16.  Feeding the same numbers, de-interleaved into the reference's per-chain order, through the
    oracle's `_update_fantasy_data` must give the same binary states."""
    rng = np.random.default_rng(9)
    V, H, n = 20, 12, 7
    m, r = _pair(rng, V, H)
    v0 = rng.integers(0, 2, (n, V)).astype(np.float64)
    rs = np.random.RandomState(123)
    U_h, U_v = np.empty((k, n, H)), np.empty((k, n, V))
    v = v0.copy()
    hs = None
    for s in range(k):
        state = rs.get_state()
        U_h[s] = rs.uniform(size=(n, H))
        U_v[s] = rs.uniform(size=(n, V))
        rs.set_state(state)
        hs = r._sample_hiddens(v, rs)
        v = r._sample_visibles(hs, rs).astype(np.float64)
    chains = [qo.FantasyData(v0[i].copy(), np.zeros(H)) for i in range(n)]
    qo._update_fantasy_data(m, chains, k, qo.Stream(qo.serial_uniforms_pcd(U_h, U_v, None)))
    assert np.array_equal(np.stack([c.v for c in chains]), v)
    assert np.array_equal(np.stack([c.h for c in chains]), hs.astype(np.float64))   # the stored h is the sample drawn from the previous v
