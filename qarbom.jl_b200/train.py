"""`train!` for CD and PCD with the reference's keyword arguments, log format and CSV schema
(src/training/train_cd.jl:45-200, src/training/train_pcd.jl:133-345); the per-epoch work
(`contrastive_divergence!`, `persistent_contrastive_divergence!`, `evaluate`) runs on the
B200 through the C ABI.  New keyword arguments default to the reference's behaviour."""
from __future__ import annotations

import csv
import warnings
from typing import Optional, Sequence

import numpy as np

from .device import DeviceRBM
from .rbm import AbstractRBM, copy_rbm, copy_rbm_into


class TrainingMethod: ...
class CD(TrainingMethod): ...
class PCD(TrainingMethod): ...
class FastPCD(TrainingMethod): ...
class EvaluationMethod: ...
class MeanSquaredError(EvaluationMethod): ...
class Accuracy(EvaluationMethod): ...
class CrossEntropy(EvaluationMethod): ...
class FreeEnergy(EvaluationMethod): ...   # new metric (no reference counterpart)


def _diverged(history: Sequence[float]) -> bool:
    """src/evaluation.jl:132-154 (MeanSquaredError, CrossEntropy: same rule; NaN compares false, as in Julia)."""
    if len(history) == 1:
        return False
    return history[-1] > history[-2] or history[-1] > min(history)


def _diverged_accuracy(history: Sequence[float]) -> bool:
    """src/evaluation.jl:156-168 (Accuracy)."""
    if len(history) == 1:
        return False
    if history[-1] < history[-2]:
        print(f"Accuracy decreased from {history[-2]} to {history[-1]}")
        return True
    if history[-1] < max(history):
        print(f"Accuracy {history[-1]} is not the maximum value {max(history)}")
        return True
    return False


def _log_epoch(epoch, t_sample, t_gibbs, t_update, total_t):
    """src/utils.jl:28-43, byte-compatible."""
    bar = "|------------------------------------------------------------------|"
    print(bar)
    print("| Epoch | Time (Sample) | Time (Gibbs) | Time (Update) | Total     |")
    print(bar)
    print("| %5d | %13.4f | %12.4f | %13.4f | %9.4f |" % (epoch, t_sample, t_gibbs, t_update, total_t))
    print(bar)


def _log_finish(n_epochs, ts, tg, tu):
    """src/utils.jl:68-74."""
    print(f"QARBoM.jl finished training after {n_epochs} epochs.")
    print(f"Total time spent sampling: {ts}")
    print(f"Total time spent in Gibbs sampling: {tg}")
    print(f"Total time spent updating parameters: {tu}")
    print(f"Total time spent training: {ts + tg + tu}")


def train(rbm: AbstractRBM, x_train, method, label_train=None, *, n_epochs: int, learning_rate: Sequence[float],
          gibbs_steps: Optional[int] = None, batch_size: Optional[int] = None,
          label_learning_rate: Optional[Sequence[float]] = None, fast_learning_rate: Optional[float] = None,
          fast_label_learning_rate: float = 0.1, metrics=None,
          early_stopping: bool = False, store_best_rbm: bool = True, patience: int = 10,
          stopping_metric=None, x_test_dataset=None, y_test_dataset=None, file_path: Optional[str] = None,
          # ---- extensions.  Their defaults reproduce the reference's SEMANTICS; the arithmetic defaults to the bf16 production
          # mode (increments within 2e-2 of Float64).  precision="fp32" is the validation mode (1e-5 of Float64).
          momentum: float = 0.0, weight_decay: float = 0.0, precision: str = "bf16", seed: int = 0, device: int = 0,
          n_chains: Optional[int] = None,
          verbose: bool = True, chains_init=None) -> dict:
    is_fast = method is FastPCD or isinstance(method, FastPCD)
    is_pcd = is_fast or method is PCD or isinstance(method, PCD)
    clf = rbm.n_classifiers > 0
    if is_fast and fast_learning_rate is None:
        if not clf:
            raise TypeError("train!(..., FastPCD) requires fast_learning_rate")   # keyword without default, train_fast_pcd.jl:177
        fast_learning_rate = 0.1                                                  # classifier default, train_fast_pcd.jl:290
    if clf and label_train is None:
        raise ValueError("label_train is required for classifier models")
    if clf and label_learning_rate is None:
        raise TypeError("train!(::RBMClassifiers, ...) requires label_learning_rate")   # train_cd.jl:129
    if metrics is None:          # defaults: [MeanSquaredError] (train_cd.jl:52) / [Accuracy] (train_cd.jl:130)
        metrics = (Accuracy,) if clf else (MeanSquaredError,)
    if stopping_metric is None:
        stopping_metric = Accuracy if clf else MeanSquaredError
    if is_pcd and batch_size is None:
        raise TypeError("train!(..., PCD) requires batch_size")   # keyword without default, train_pcd.jl:139
    if gibbs_steps is None:
        gibbs_steps = 1 if is_pcd else 3                         # train_pcd.jl:138 / train_cd.jl:50
    B = int(batch_size) if batch_size is not None else 1        # CD default = the reference's online update
    if file_path is None:
        file_path = ("fast_pcd" if is_fast else "pcd" if is_pcd else "cd") + ("_classifier" if clf else "") + "_metrics.csv"
    n = len(x_train)
    if n % B:
        # `_set_mini_batches`, src/utils.jl:13-26
        warnings.warn(f"The last batch size is not equal to the other batches. Will dismiss {n % B} samples.")
    if len(learning_rate) < n_epochs:
        raise IndexError("learning_rate needs one entry per epoch")       # learning_rate[epoch], train_cd.jl:77

    names = {MeanSquaredError: "mse", FreeEnergy: "free_energy", Accuracy: "accuracy", CrossEntropy: "cross_entropy"}
    keys = [names[m] for m in metrics]
    nc = B if n_chains is None else int(n_chains)                  # reference: n_chains == batch_size (train_pcd.jl:162)
    dev = DeviceRBM(rbm, max_batch=max(B, nc, min(n, 4096)), precision=precision, device=device, seed=seed)
    try:
        dev.set_optimizer(momentum, weight_decay)
        dev.set_data(x_train, label_train)

        use_test = x_test_dataset is not None and (not clf or y_test_dataset is not None)  # train_cd.jl:143-147

        def evaluate():
            out = {}
            for m in metrics:
                xs = x_test_dataset if use_test else None
                if m is MeanSquaredError:
                    out[names[m]] = dev.eval_mse(xs)
                elif m is Accuracy:
                    out[names[m]] = dev.eval_accuracy(xs, y_test_dataset if use_test else label_train)
                elif m is CrossEntropy:
                    out[names[m]] = dev.eval_cross_entropy(xs, y_test_dataset if use_test else label_train)
                else:
                    out[names[m]] = dev.eval_free_energy(xs)
            return out

        dev.snapshot_params()                                             # best_rbm = copy_rbm(rbm), kept in HBM
        initial = evaluate()                                              # initial_evaluation
        hist = {k: [] for k in keys}
        initial_patience = patience
        tot = [0.0, 0.0, 0.0]
        if is_pcd:
            if verbose:
                print("Setting mini-batches")
            if chains_init is not None:
                dev.init_chains(nc, *chains_init)
            else:
                dev.init_chains(nc)                                        # _init_fantasy_data(rbm, batch_size)
            if verbose:
                print("Starting training")
        stop_key = names.get(stopping_metric, keys[0])
        epochs_run = n_epochs
        for epoch in range(1, n_epochs + 1):
            lr = float(learning_rate[epoch - 1])
            llr = float(label_learning_rate[epoch - 1]) if clf else 0.0
            if is_fast:
                t = dev.fast_pcd_epoch(B, gibbs_steps, lr, float(fast_learning_rate), llr, float(fast_label_learning_rate))
            else:
                t = dev.pcd_epoch(B, gibbs_steps, lr, llr) if is_pcd else dev.cd_epoch(B, gibbs_steps, lr, llr)
            for i in range(3):
                tot[i] += t[i]
            for k, v in evaluate().items():
                hist[k].append(v)
            diverged = _diverged_accuracy(hist[stop_key]) if stop_key == "accuracy" else _diverged(hist[stop_key])
            if diverged:
                if early_stopping:
                    if patience == 0:
                        print(f"Early stopping at epoch {epoch}")
                        epochs_run = epoch
                        break
                    patience -= 1
            else:
                patience = initial_patience
                if store_best_rbm:
                    dev.snapshot_params()                                 # copy_rbm!(rbm, best_rbm)
            if verbose:
                _log_epoch(epoch, t[0], t[1], t[2], sum(tot))
                for k in keys:
                    print(f"{k}: {hist[k][-1]}")
        if store_best_rbm:
            dev.restore_params()                                          # copy_rbm!(best_rbm, rbm)
        dev.pull_params()
        merged = {k: [initial[k]] + hist[k] for k in keys}                # merge_metrics
        if file_path:
            # CSV.write(file_path, DataFrame(metrics_dict)), train_cd.jl:113-115: a DataFrame built from a Dict sorts its columns by name
            cols = sorted(keys)
            with open(file_path, "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(cols)
                for row in zip(*[merged[k] for k in cols]):
                    w.writerow(row)
        if verbose:
            _log_finish(epochs_run, *tot)
        return merged
    finally:
        dev.close()


def reconstruct(rbm: AbstractRBM, v, precision: str = "fp32", device: int = 0) -> np.ndarray:
    """QARBoM.reconstruct(rbm, v) (src/rbm.jl:220-224) on the device."""
    v = np.asarray(v, dtype=np.float64)
    with DeviceRBM(rbm, max_batch=max(1, min(4096, v.reshape(-1, rbm.n_visible).shape[0])), precision=precision,
                   device=device) as dev:
        out = dev.reconstruct(v)
    return out.reshape(v.shape)
