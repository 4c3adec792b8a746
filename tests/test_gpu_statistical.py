"""Production RNG mode: learning curves under the device's Philox stream must agree statistically with
the CPU oracle driven by an independent generator (SURVEY.md section 8(c) iv): same model, data and
hyper-parameters, 10 seeds each, final reconstruction error and free energy compared through their
seed-to-seed spread."""
import numpy as np
import pytest

from oracle import qarbom_oracle as qo
from tests.helpers import f32_round, to_host_rbm

pytestmark = pytest.mark.gpu


def _dataset(rng, n, V):
    protos = (rng.random((4, V)) < 0.5).astype(np.float64)
    X = protos[rng.integers(0, 4, n)]
    flip = rng.random((n, V)) < 0.05
    return np.where(flip, 1.0 - X, X)


@pytest.mark.parametrize("method,prec", [("PCD", "bf16"), ("PCD", "fp32"), ("CD", "bf16")])
def test_learning_curves_match_oracle_statistically(method, prec):
    import qarbom_b200 as q
    q.build()
    rng = np.random.default_rng(4)
    V, H, n, B, epochs, lr, seeds = 36, 16, 256, 32, 6, 0.1, 10
    X = _dataset(rng, n, V)
    W0 = f32_round(0.05 * rng.standard_normal((V, H)))
    k = 1

    def oracle_run(seed):
        m = qo.make_model(V, H, W=W0)
        st = qo.NumpyStream(1000 + seed)
        if method == "PCD":
            chains = qo._init_fantasy_data(m, B, st)
            for _ in range(epochs):
                qo.persistent_contrastive_divergence(m, X, qo._set_mini_batches(n, B), chains, steps=k, learning_rate=lr, stream=st)
        else:
            for _ in range(epochs):
                qo.contrastive_divergence(m, X, steps=k, learning_rate=lr, stream=st, batch_size=B)
        return qo.evaluate_mse(m, X), qo.mean_free_energy(m, X)

    def device_run(seed):
        rbm = q.RBM(V, H, W=W0)
        with q.DeviceRBM(rbm, max_batch=B, precision=prec, seed=seed) as dev:
            dev.set_data(X)
            if method == "PCD":
                dev.init_chains(B)
            for _ in range(epochs):
                (dev.pcd_epoch if method == "PCD" else dev.cd_epoch)(B, k, lr, 0.0, timed=False)
            return dev.eval_mse(), dev.eval_free_energy()

    o = np.array([oracle_run(s) for s in range(seeds)])
    d = np.array([device_run(s) for s in range(seeds)])
    mse0 = qo.evaluate_mse(qo.make_model(V, H, W=W0), X)
    assert o[:, 0].mean() < 0.8 * mse0 and d[:, 0].mean() < 0.8 * mse0          # both actually learn
    for j, name in enumerate(("mse", "free_energy")):
        se = np.sqrt(o[:, j].var(ddof=1) / seeds + d[:, j].var(ddof=1) / seeds)
        diff = abs(o[:, j].mean() - d[:, j].mean())
        assert diff < 4 * se + 0.02 * abs(o[:, j].mean()), (name, o[:, j].mean(), d[:, j].mean(), se)
