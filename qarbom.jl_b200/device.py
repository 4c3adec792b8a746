"""`DeviceRBM`: one C-ABI handle bound to one host-side model (the Python stand-in for the
Julia shim, which cannot be executed in this image -- see INTEGRATION.md)."""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import QbConfig, QbError, as_rows, check, lib, ptr


class DeviceRBM:
    def __init__(self, rbm, max_batch: int, precision: str = "bf16", device: int = 0, rank: int = 0, world: int = 1,
                 seed: int = 0):
        self.rbm = rbm
        self.V, self.H, self.L = rbm.n_visible, rbm.n_hidden, getattr(rbm, "n_classifiers", 0)
        cfg = QbConfig(self.V, self.H, self.L,
                       _ffi.QB_VISIBLE_GAUSSIAN if rbm.gaussian else _ffi.QB_VISIBLE_BERNOULLI,
                       {"fp32": _ffi.QB_PREC_FP32, "bf16": _ffi.QB_PREC_BF16, "fp32x3": _ffi.QB_PREC_FP32X3}[precision], int(max_batch), device, rank,
                       world, 0)
        self._h = ctypes.c_void_p()
        check(lib().qb_create(ctypes.byref(self._h), ctypes.byref(cfg)))
        self.world, self.rank = world, rank
        self.push_params()
        self.seed(seed)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().qb_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ parameters
    def push_params(self):
        r = self.rbm
        W = np.asfortranarray(r.W, dtype=np.float64)  # Julia Matrix{Float64}: column-major V x H
        U = np.asfortranarray(r.U, dtype=np.float64) if self.L else None
        c = np.ascontiguousarray(r.c, dtype=np.float64) if self.L else None
        a = np.ascontiguousarray(r.a, dtype=np.float64)
        b = np.ascontiguousarray(r.b, dtype=np.float64)
        check(lib().qb_set_params(self._h, ptr(W), ptr(a), ptr(b), ptr(U), ptr(c)))

    def pull_params(self):
        r = self.rbm
        W = np.zeros((self.V, self.H), order="F")
        a, b = np.zeros(self.V), np.zeros(self.H)
        U = np.zeros((self.L, self.H), order="F") if self.L else None
        c = np.zeros(self.L) if self.L else None
        check(lib().qb_get_params(self._h, ptr(W), ptr(a), ptr(b), ptr(U), ptr(c)))
        r.W[...] = W
        r.a[...] = a
        r.b[...] = b
        if self.L:
            r.U[...] = U
            r.c[...] = c

    def snapshot_params(self):
        """best_rbm = copy_rbm(rbm), on the device."""
        check(lib().qb_snapshot_params(self._h))

    def restore_params(self):
        """copy_rbm!(best_rbm, rbm), on the device."""
        check(lib().qb_restore_params(self._h))

    def set_optimizer(self, momentum: float = 0.0, weight_decay: float = 0.0):
        check(lib().qb_set_optimizer(self._h, momentum, weight_decay))

    def seed(self, seed: int):
        check(lib().qb_seed(self._h, ctypes.c_uint64(seed)))

    def set_option(self, key: str, value: float):
        check(lib().qb_set_option(self._h, key.encode(), float(value)))

    # ------------------------------------------------------------------ data / chains
    def set_data(self, x, y=None):
        X, xdt = as_rows(x, self.V)
        Y, ydt = as_rows(y, self.L) if self.L else (None, 0)
        check(lib().qb_set_data(self._h, ptr(X), xdt, ptr(Y), ydt, X.shape[0]))
        self.n_rows = X.shape[0]

    def init_chains(self, n: int, v=None, h=None, y=None):
        v = None if v is None else np.ascontiguousarray(v, dtype=np.float64)
        h = None if h is None else np.ascontiguousarray(h, dtype=np.float64)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        check(lib().qb_init_chains(self._h, n, ptr(v), ptr(h), ptr(y)))
        self.n_chains = n

    def get_chains(self):
        n = self.n_chains
        v, h = np.zeros((n, self.V)), np.zeros((n, self.H))
        y = np.zeros((n, self.L)) if self.L else None
        check(lib().qb_get_chains(self._h, ptr(v), ptr(h), ptr(y)))
        return v, h, y

    def get_chain_rows(self, row0: int, n: int):
        v, h = np.zeros((n, self.V)), np.zeros((n, self.H))
        y = np.zeros((n, self.L)) if self.L else None
        check(lib().qb_get_chain_rows(self._h, row0, n, ptr(v), ptr(h), ptr(y)))
        return v, h, y

    def inject_streams(self, U_h, U_v=None, U_y=None, Z_v=None):
        U_h = np.ascontiguousarray(U_h, dtype=np.float64)
        k, B = U_h.shape[0], U_h.shape[1]
        U_v = None if U_v is None else np.ascontiguousarray(U_v, dtype=np.float64)
        U_y = None if U_y is None else np.ascontiguousarray(U_y, dtype=np.float64)
        Z_v = None if Z_v is None else np.ascontiguousarray(Z_v, dtype=np.float64)
        check(lib().qb_inject_streams(self._h, ptr(U_h), ptr(U_v), ptr(U_y), ptr(Z_v), k, B))

    # ------------------------------------------------------------------ training steps
    def pcd_step(self, batch_index: int, B: int, k: int, lr: float, llr: float = 0.0):
        check(lib().qb_pcd_step(self._h, batch_index, B, k, lr, llr))

    def qsampling_step(self, batch_index: int, B: int, v_model, h_model, lr: float, y_model=None, llr: float = 0.0):
        """One mini-batch of persistent_qubo! (src/training/train_quantum_pcd.jl:10-37) with the sampler's
        (v_model, h_model[, label_model]) supplied by the caller: positive phase, gradient and update on the GPU."""
        self.init_chains(1, np.asarray(v_model, dtype=np.float64).reshape(1, -1),
                         np.asarray(h_model, dtype=np.float64).reshape(1, -1),
                         None if y_model is None else np.asarray(y_model, dtype=np.float64).reshape(1, -1))
        check(lib().qb_pcd_step(self._h, batch_index, B, 0, lr, llr))

    def cd_step(self, batch_index: int, B: int, k: int, lr: float, llr: float = 0.0):
        check(lib().qb_cd_step(self._h, batch_index, B, k, lr, llr))

    def pcd_epoch(self, B: int, k: int, lr: float, llr: float = 0.0, timed: bool = True):
        t = (ctypes.c_double * 3)()
        check(lib().qb_pcd_epoch(self._h, B, k, lr, llr, t if timed else None))
        return tuple(t)

    def fast_pcd_epoch(self, B: int, k: int, lr: float, fast_lr: float, llr: float = 0.0, fast_llr: float = 0.0,
                       timed: bool = True):
        """fast_persistent_contrastive_divergence! for one epoch (src/training/train_fast_pcd.jl:1-127)."""
        t = (ctypes.c_double * 3)()
        check(lib().qb_fast_pcd_epoch(self._h, B, k, lr, fast_lr, llr, fast_llr, t if timed else None))
        return tuple(t)

    def train_steps(self, method: str, first_batch: int, n_steps: int, B: int, k: int, lr: float, llr: float = 0.0):
        """n_steps consecutive mini-batch steps on resident batches first_batch, first_batch + 1, ... (qb_train_steps)."""
        check(lib().qb_train_steps(self._h, 1 if method.upper() == "PCD" else 0, first_batch, n_steps, B, k, lr, llr))

    METHODS = {"CD": 0, "PCD": 1, "FASTPCD": 2}
    METRICS = {"mse": 0, "accuracy": 1, "cross_entropy": 2, "free_energy": 3}

    def train_epochs(self, method: str, n_epochs: int, B: int, k: int, learning_rate, label_learning_rate=None, *,
                     metrics=("mse",), stopping_metric=None, early_stopping=False, store_best_rbm=True, patience=10,
                     fast_learning_rate=0.1, fast_label_learning_rate=0.1, x_eval=None, y_eval=None, on_epoch=None):
        """train!'s whole epoch loop inside the library (qb_train): initial evaluation, epochs, evaluate, _diverged, patience,
        best-model snapshot / restore in HBM.  Returns (history[(epochs_run + 1)][n_metrics], (t_sample, t_gibbs, t_update),
        epochs_run); on_epoch(epoch, metrics_row, times, flags) is called after every epoch."""
        lr = np.ascontiguousarray(learning_rate, dtype=np.float64)
        llr = None if label_learning_rate is None else np.ascontiguousarray(label_learning_rate, dtype=np.float64)
        ids = np.ascontiguousarray([self.METRICS[m] for m in metrics], dtype=np.int32)
        stop = self.METRICS[stopping_metric if stopping_metric is not None else metrics[0]]
        X, xdt = as_rows(x_eval, self.V) if x_eval is not None else (None, 0)
        Y, ydt = as_rows(y_eval, self.L) if y_eval is not None else (None, 0)
        hist = np.full((n_epochs + 1, len(ids)), np.nan)
        times = np.zeros(3)
        ran = ctypes.c_int32(0)

        def _cb(user, epoch, mrow, trow, flags):
            on_epoch(int(epoch), [mrow[i] for i in range(len(ids))], [trow[i] for i in range(3)], int(flags))

        cb = _ffi.ON_EPOCH(_cb) if on_epoch is not None else ctypes.cast(None, _ffi.ON_EPOCH)
        cfg = _ffi.QbTrainConfig(self.METHODS[method.upper()], n_epochs, B, k, lr.ctypes.data, None if llr is None else llr.ctypes.data,
                                 fast_learning_rate, fast_label_learning_rate, len(ids), ids.ctypes.data, stop,
                                 int(early_stopping), int(store_best_rbm), patience,
                                 None if X is None else X.ctypes.data, xdt, 0 if X is None else X.shape[0],
                                 None if Y is None else Y.ctypes.data, ydt, cb, None)
        check(lib().qb_train(self._h, ctypes.byref(cfg), ptr(hist), ptr(times), ctypes.byref(ran)))
        return hist[:ran.value + 1], tuple(times), ran.value

    def cd_epoch(self, B: int, k: int, lr: float, llr: float = 0.0, timed: bool = True):
        t = (ctypes.c_double * 3)()
        check(lib().qb_cd_epoch(self._h, B, k, lr, llr, t if timed else None))
        return tuple(t)

    def step_host(self, method: str, x, y, k: int, lr: float, llr: float = 0.0, want_recon: bool = True,
                  next_x=None, next_y=None) -> float:
        """One step on a host mini-batch.  `next_x`/`next_y` (the batch of the following call) are uploaded on a
        copy stream while this step computes; pass the same arrays as x/y of the next call."""
        X, xdt = as_rows(x, self.V)
        Y, ydt = as_rows(y, self.L) if self.L else (None, 0)
        if next_x is not None:
            NX, nxdt = as_rows(next_x, self.V)
            NY, nydt = as_rows(next_y, self.L) if self.L else (None, 0)
            check(lib().qb_set_next_host_batch(self._h, ptr(NX), nxdt, ptr(NY), nydt))
            self._keepalive = (NX, NY)  # borrowed by the library until the consuming call returns
        out = ctypes.c_double(0.0)
        fn = lib().qb_pcd_step_host if method.upper() == "PCD" else lib().qb_cd_step_host
        check(fn(self._h, ptr(X), xdt, ptr(Y), ydt, X.shape[0], k, lr, llr, ctypes.byref(out) if want_recon else None))
        return out.value

    # ------------------------------------------------------------------ evaluation
    def eval_mse(self, x=None, Z=None) -> float:
        out = ctypes.c_double(0.0)
        Z = None if Z is None else np.ascontiguousarray(Z, dtype=np.float64)
        if x is None:
            check(lib().qb_eval_mse(self._h, None, 0, 0, ptr(Z), ctypes.byref(out)))
        else:
            X, xdt = as_rows(x, self.V)
            check(lib().qb_eval_mse(self._h, ptr(X), xdt, X.shape[0], ptr(Z), ctypes.byref(out)))
        return out.value

    def eval_free_energy(self, x=None, y=None) -> float:
        out = ctypes.c_double(0.0)
        if x is None:
            check(lib().qb_eval_free_energy(self._h, None, 0, None, 0, 0, ctypes.byref(out)))
        else:
            X, xdt = as_rows(x, self.V)
            Y, ydt = as_rows(y, self.L) if (self.L and y is not None) else (None, 0)
            check(lib().qb_eval_free_energy(self._h, ptr(X), xdt, ptr(Y), ydt, X.shape[0], ctypes.byref(out)))
        return out.value

    def reconstruct(self, v, Z=None) -> np.ndarray:
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1, self.V))
        Z = None if Z is None else np.ascontiguousarray(Z, dtype=np.float64)
        out = np.zeros_like(v)
        check(lib().qb_reconstruct(self._h, ptr(v), v.shape[0], ptr(Z), ptr(out)))
        return out

    def conditional_prob_h(self, v, y=None) -> np.ndarray:
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1, self.V))
        y = None if y is None else np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1, self.L))
        out = np.zeros((v.shape[0], self.H))
        check(lib().qb_conditional_prob_h(self._h, ptr(v), ptr(y), v.shape[0], ptr(out)))
        return out

    def sample_hidden(self, v, y=None, U=None) -> np.ndarray:
        """gibbs_sample_hidden (src/gibbs.jl:3-6,29-32) per row; U = the caller's uniforms [rows][H], None = Philox."""
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1, self.V))
        y = None if y is None else np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1, self.L))
        U = None if U is None else np.ascontiguousarray(np.asarray(U, dtype=np.float64).reshape(-1, self.H))
        out = np.zeros((v.shape[0], self.H))
        check(lib().qb_sample_hidden(self._h, ptr(v), ptr(y), v.shape[0], ptr(U), ptr(out)))
        return out

    def sample_visible(self, h, R=None, Ry=None):
        """gibbs_sample_visible (+ gibbs_sample_label for classifiers) per row (src/gibbs.jl:13-20,46-49).  R: uniforms
        (standard normals for Gaussian visibles) [rows][V], Ry: label uniforms [rows][L]; None = Philox.
        Returns v, or (v, y) for classifiers."""
        h = np.ascontiguousarray(np.asarray(h, dtype=np.float64).reshape(-1, self.H))
        R = None if R is None else np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(-1, self.V))
        Ry = None if Ry is None else np.ascontiguousarray(np.asarray(Ry, dtype=np.float64).reshape(-1, self.L))
        out_v = np.zeros((h.shape[0], self.V))
        out_y = np.zeros((h.shape[0], self.L)) if self.L else None
        check(lib().qb_sample_visible(self._h, ptr(h), h.shape[0], ptr(R), ptr(Ry), ptr(out_v), ptr(out_y)))
        return (out_v, out_y) if self.L else out_v

    def conditional_prob_y_given_v(self, x=None) -> np.ndarray:
        """softmax(c + U sigma(b + W'v)) per row (src/rbm.jl:191-194); x=None: the resident dataset."""
        if x is None:
            out = np.zeros((self.n_rows, self.L))
            check(lib().qb_conditional_prob_y_given_v(self._h, None, 0, 0, ptr(out)))
            return out
        X, xdt = as_rows(x, self.V)
        out = np.zeros((X.shape[0], self.L))
        check(lib().qb_conditional_prob_y_given_v(self._h, ptr(X), xdt, X.shape[0], ptr(out)))
        return out

    def classify(self, x=None) -> np.ndarray:
        """classify (src/rbm.jl:226-229): 1 where the label probability equals the row maximum."""
        p = self.conditional_prob_y_given_v(x)
        return (p == p.max(axis=1, keepdims=True)).astype(np.int64)

    def eval_cross_entropy(self, x, y) -> float:
        """_evaluate(CrossEntropy) through the classifier `evaluate` (src/evaluation.jl:22-26,28-50), literally:
        the reference feeds the ROUNDED prediction (1 at the row maximum, else 0) into
        sum(y log p + (1 - y) log(1 - p)) / N, so every term is 0 * -Inf or -Inf and the metric is NaN (or -Inf)
        for any data -- reproduced, not repaired."""
        pred = self.classify(x).astype(np.float64)
        y = np.asarray(y, dtype=np.float64).reshape(pred.shape)
        with np.errstate(divide="ignore", invalid="ignore"):
            per_sample = np.sum(y * np.log(pred) + (1.0 - y) * np.log(1.0 - pred), axis=1)
            return float(np.sum(per_sample / pred.shape[0]))

    def eval_accuracy(self, x, y) -> float:
        """_evaluate(Accuracy) through the classifier `evaluate` (src/evaluation.jl:9-14,28-50)."""
        pred = self.classify(x)
        y = np.rint(np.asarray(y, dtype=np.float64).reshape(pred.shape)).astype(np.int64)
        return float(np.mean(np.all(pred == y, axis=1)))

    # ------------------------------------------------------------------ multi-GPU / counters
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        check(lib().qb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes):
        # the library dlopens NCCL itself: point it at the nvidia-nccl wheel (the copy torch.distributed uses, and the one whose
        # device-API headers the NVSwitch-multicast exchange was built against) unless the caller chose a library
        if not os.environ.get("QB_NCCL_LIB"):
            from .build import nccl_runtime_library
            path = nccl_runtime_library()
            if path:
                os.environ["QB_NCCL_LIB"] = path
        buf = ctypes.create_string_buffer(uid, 128)
        check(lib().qb_comm_init(self._h, buf))

    def enable_nvls_exchange(self) -> str:
        """COLLECTIVE (every rank, after comm_init, between steps): switch the gradient exchange to the fused kernel over
        NVSwitch multicast (multimem.ld_reduce / multimem.st).  Returns "" on success, else the reason why multicast is
        unavailable -- the NCCL reduce-scatter / all-gather exchange stays in place then."""
        try:
            self.set_option("nvls_exchange", 1)
            return ""
        except QbError as e:
            if e.code != -5:   # QB_E_UNSUPPORTED
                raise
            return str(e)

    def p2p_export(self) -> bytes:
        buf = ctypes.create_string_buffer(320)
        check(lib().qb_p2p_export(self._h, buf))
        return buf.raw

    def p2p_import(self, all_handles: bytes):
        buf = ctypes.create_string_buffer(all_handles, len(all_handles))
        check(lib().qb_p2p_import(self._h, buf))

    def p2p_timing(self):
        a, b = ctypes.c_double(0), ctypes.c_double(0)
        check(lib().qb_get_p2p_timing(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def counters(self):
        """(kernel launches, GEMM launches) since the last reset."""
        return self.counters_full()[:2]

    def counters_full(self):
        a, b, t = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        f, ms, tf = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
        check(lib().qb_get_counters(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(f), ctypes.byref(t),
                                    ctypes.byref(ms), ctypes.byref(tf)))
        return a.value, b.value, f.value, t.value, ms.value, tf.value

    GEMM_KINDS = ("generic", "grad", "fast_sample", "fast_prob", "fast_sample_prob", "fast_gauss", "fast_label", "chain", "epoch", "grad_x3", "fast_sqerr")

    def gemm_histogram(self) -> dict:
        """{(kernel, tile width, kind name): launches} since the last counter reset."""
        nk = len(self.GEMM_KINDS)
        n = 2 * 4 * nk + 1
        c = (ctypes.c_int64 * n)()
        check(lib().qb_get_gemm_histogram(self._h, c, n))
        out = {}
        for pair in range(2):
            for b in range(4):
                for k in range(nk):
                    v = c[(pair * 4 + b) * nk + k]
                    if v:
                        out[("2cta" if pair else "1cta", 32 << b, self.GEMM_KINDS[k])] = int(v)
        if c[n - 1]:
            out[("simt", 64, "fp32")] = int(c[n - 1])
        return out

    def issued_flops(self) -> float:
        f = ctypes.c_double(0)
        check(lib().qb_get_issued_flops(self._h, ctypes.byref(f)))
        return f.value

    def update_timing(self):
        n, ms = ctypes.c_int64(0), ctypes.c_double(0)
        check(lib().qb_get_update_timing(self._h, ctypes.byref(n), ctypes.byref(ms)))
        return n.value, ms.value

    def reset_counters(self):
        check(lib().qb_reset_counters(self._h))

    def sync(self):
        check(lib().qb_sync(self._h))

    def timer_start(self):
        check(lib().qb_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = ctypes.c_double(0)
        check(lib().qb_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value
