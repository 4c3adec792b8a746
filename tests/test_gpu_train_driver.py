"""qb_train -- train!'s epoch loop inside the library (initial evaluation, epochs, evaluate, _diverged, patience, best-model
snapshot / restore in HBM) -- against the verified host loop of qarbom.jl_b200/train.py, which issues the same
library calls one by one (src/training/train_cd.jl:60-115 and its PCD / FastPCD / classifier variants).  Same seed, same
calls: metric histories and final parameters must agree.  Run with -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NAMES = {"mse": "MeanSquaredError", "accuracy": "Accuracy", "cross_entropy": "CrossEntropy", "free_energy": "FreeEnergy"}


@pytest.fixture(scope="module")
def q():
    import qarbom_b200
    qarbom_b200.build()
    return qarbom_b200


def _data(rng, n, V, L):
    proto = (rng.random((max(L, 2), V)) < 0.4).astype(np.int64)
    lab = rng.integers(0, max(L, 2), n)
    X = np.where(rng.random((n, V)) < 0.08, 1 - proto[lab], proto[lab])
    Y = np.eye(L, dtype=np.int64)[lab] if L else None
    return X, Y


CASES = [  # method, V, H, L, B, k, epochs, metrics, stopping, early_stopping, patience, with test set
    ("PCD", 20, 12, 0, 8, 1, 6, ("mse",), "mse", False, 10, False),
    ("PCD", 20, 12, 0, 8, 2, 12, ("mse", "free_energy"), "mse", True, 1, True),          # early stopping on a test set
    ("CD", 16, 10, 0, 1, 2, 4, ("mse",), "mse", False, 10, False),                        # the reference's online CD
    ("CD", 24, 14, 3, 1, 1, 5, ("accuracy", "cross_entropy"), "accuracy", True, 2, False),   # classifier defaults (+ the literal CrossEntropy)
    ("PCD", 24, 14, 3, 6, 1, 5, ("accuracy", "mse"), "accuracy", False, 10, True),
    ("FastPCD", 24, 14, 3, 6, 1, 4, ("accuracy",), "accuracy", False, 10, False),
]


@pytest.mark.parametrize("method,V,H,L,B,k,epochs,metrics,stopping,early,patience,with_test", CASES)
def test_device_epoch_driver_equals_host_loop(q, tmp_path, method, V, H, L, B, k, epochs, metrics, stopping, early, patience, with_test):
    rng = np.random.default_rng(V + H + L + B + epochs)
    X, Y = _data(rng, 48, V, L)
    Xt, Yt = _data(rng, 18, V, L) if with_test else (None, None)
    W0 = 0.1 * rng.standard_normal((V, H))
    U0 = 0.1 * rng.standard_normal((L, H)) if L else None
    lr = [0.05 / (e + 1) ** 0.5 for e in range(epochs)]
    llr = [0.03] * epochs if L else None

    def model():
        return q.RBMClassifier(V, H, L, W=W0, U=U0) if L else q.RBM(V, H, W=W0)

    # ---- the host loop (train.py)
    ra = model()
    kw = dict(n_epochs=epochs, learning_rate=lr, gibbs_steps=k, metrics=[getattr(q, NAMES[m]) for m in metrics],
              stopping_metric=getattr(q, NAMES[stopping]), early_stopping=early, patience=patience, precision="fp32", seed=5,
              file_path=str(tmp_path / "a.csv"), verbose=False, x_test_dataset=Xt)
    if method != "CD":
        kw["batch_size"] = B
    if L:
        kw.update(label_learning_rate=llr, y_test_dataset=Yt)
    if method == "FastPCD":
        kw.update(fast_learning_rate=0.05, fast_label_learning_rate=0.04)
    ha = q.train(ra, X.tolist(), getattr(q, method), Y.tolist() if L else None, **kw)

    # ---- the same loop inside the library
    rb = model()
    seen = []
    with q.DeviceRBM(rb, max_batch=max(B, min(len(X), 4096)), precision="fp32", seed=5) as dev:
        dev.set_data(X, Y)
        use_test = Xt is not None and (not L or Yt is not None)
        hist, times, ran = dev.train_epochs(method, epochs, B, k, lr, llr, metrics=metrics, stopping_metric=stopping,
                                            early_stopping=early, store_best_rbm=True, patience=patience,
                                            fast_learning_rate=0.05, fast_label_learning_rate=0.04,
                                            x_eval=Xt if use_test else None, y_eval=(Yt if use_test else Y) if L else None,
                                            on_epoch=lambda e, m, t, f: seen.append((e, list(m), f)))
        dev.pull_params()
    assert ran == len(ha[metrics[0]]) - 1 and hist.shape == (ran + 1, len(metrics))
    for j, name in enumerate(metrics):
        a, b = np.asarray(ha[name], dtype=np.float64), hist[:, j]
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=1e-9, atol=1e-12), (name, a, b)
    assert [e for e, _, _ in seen] == list(range(1, ran + 1)) and all(np.allclose(np.nan_to_num(m), np.nan_to_num(hist[e])) for e, m, _ in seen)
    if early and ran < epochs:
        assert seen[-1][2] & 2          # QB_EPOCH_EARLY_STOP on the last epoch
    assert all(t >= 0 for t in times) and sum(times) > 0
    for a, b in ((ra.W, rb.W), (ra.a, rb.a), (ra.b, rb.b)) + (((ra.U, rb.U), (ra.c, rb.c)) if L else ()):
        assert np.max(np.abs(a - b)) < 1e-6, float(np.max(np.abs(a - b)))


def test_device_epoch_driver_argument_errors(q):
    rng = np.random.default_rng(0)
    X, _ = _data(rng, 16, 8, 0)
    with q.DeviceRBM(q.RBM(8, 4), max_batch=8, precision="fp32") as dev:
        with pytest.raises(q.QbError):
            dev.train_epochs("PCD", 2, 4, 1, [0.1, 0.1])                                    # no resident dataset
        dev.set_data(X)
        with pytest.raises(q.QbError):
            dev.train_epochs("PCD", 2, 4, 1, [0.1, 0.1], metrics=("mse",), stopping_metric="free_energy")   # KeyError in the reference
        with pytest.raises(q.QbError):
            dev.train_epochs("PCD", 2, 4, 1, [0.1, 0.1], metrics=("accuracy",))             # no label units
        with pytest.raises(q.QbError):
            dev.train_epochs("PCD", 2, 32, 1, [0.1, 0.1])                                   # batch larger than max_batch
        hist, _, ran = dev.train_epochs("PCD", 2, 4, 1, [0.1, 0.1])                         # the failed calls left the handle usable
        assert ran == 2 and hist.shape == (3, 1) and np.all(np.isfinite(hist))
        with pytest.raises(q.QbError):
            dev.set_option("nvls_exchange", 1)                                              # needs a communicator (world > 1)
        with pytest.raises(q.QbError):
            dev.set_option("chain_quad", 3)
