"""ctypes binding of include/qarbom_b200.h.  There is no fallback: a missing library or a
failing call raises."""
from __future__ import annotations

import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# QARBOM_B200_LIB: an explicit build of the same library (experiment builds of tools/, the Julia shim uses the same variable)
LIB_PATH = os.environ.get("QARBOM_B200_LIB") or os.path.join(HERE, "lib", "libqarbom_b200.so")

QB_VISIBLE_BERNOULLI, QB_VISIBLE_GAUSSIAN = 0, 1
QB_PREC_FP32, QB_PREC_BF16, QB_PREC_FP32X3 = 0, 1, 2
QB_DT_F64, QB_DT_I64, QB_DT_F32, QB_DT_U8 = 0, 1, 2, 3

# every symbol declared by include/qarbom_b200.h
SYMBOLS = [
    "qb_version", "qb_device_count", "qb_last_error", "qb_create", "qb_destroy", "qb_set_params", "qb_get_params", "qb_snapshot_params", "qb_restore_params",
    "qb_set_optimizer", "qb_set_data", "qb_init_chains", "qb_get_chains", "qb_get_chain_rows", "qb_get_gemm_histogram", "qb_get_issued_flops", "qb_seed", "qb_inject_streams",
    "qb_pcd_step", "qb_cd_step", "qb_pcd_epoch", "qb_cd_epoch", "qb_train_steps", "qb_fast_pcd_epoch", "qb_pcd_step_host", "qb_cd_step_host", "qb_set_next_host_batch", "qb_eval_mse",
    "qb_eval_free_energy", "qb_reconstruct", "qb_conditional_prob_h", "qb_sample_hidden", "qb_sample_visible", "qb_conditional_prob_y_given_v", "qb_comm_unique_id", "qb_comm_init", "qb_p2p_export", "qb_p2p_import", "qb_get_p2p_timing",
    "qb_get_counters", "qb_reset_counters", "qb_set_option", "qb_sync", "qb_timer_start", "qb_timer_stop", "qb_get_update_timing", "qb_train",
]


class QbConfig(ctypes.Structure):
    _fields_ = [("n_visible", ctypes.c_int64), ("n_hidden", ctypes.c_int64), ("n_classifiers", ctypes.c_int64),
                ("visible_kind", ctypes.c_int32), ("precision", ctypes.c_int32), ("max_batch", ctypes.c_int64),
                ("device", ctypes.c_int32), ("rank", ctypes.c_int32), ("world", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


ON_EPOCH = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.c_int32)


class QbTrainConfig(ctypes.Structure):   # mirrors `qb_train_config`
    _fields_ = [("method", ctypes.c_int32), ("n_epochs", ctypes.c_int32), ("batch_size", ctypes.c_int64), ("gibbs_steps", ctypes.c_int64),
                ("learning_rate", ctypes.c_void_p), ("label_learning_rate", ctypes.c_void_p),
                ("fast_learning_rate", ctypes.c_double), ("fast_label_learning_rate", ctypes.c_double),
                ("n_metrics", ctypes.c_int32), ("metrics", ctypes.c_void_p),
                ("stopping_metric", ctypes.c_int32), ("early_stopping", ctypes.c_int32), ("store_best_rbm", ctypes.c_int32), ("patience", ctypes.c_int32),
                ("X_eval", ctypes.c_void_p), ("x_dtype", ctypes.c_int32), ("n_eval", ctypes.c_int64),
                ("Y_eval", ctypes.c_void_p), ("y_dtype", ctypes.c_int32),
                ("on_epoch", ON_EPOCH), ("user", ctypes.c_void_p)]


class QbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"qarbom_b200 error {code}: {msg}")
        self.code = code


_lib = None
_P = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32
_f64 = ctypes.c_double


def lib():
    """Loads libqarbom_b200.so (raises if it was not built: there is no CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QbError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(qarbom_b200 has no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    L.qb_version.restype = _i32
    L.qb_last_error.restype = ctypes.c_char_p
    sig = {
        "qb_device_count": [_P],
        "qb_create": [_P, ctypes.POINTER(QbConfig)],
        "qb_destroy": [_P],
        "qb_set_params": [_P, _P, _P, _P, _P, _P],
        "qb_get_params": [_P, _P, _P, _P, _P, _P],
        "qb_snapshot_params": [_P],
        "qb_restore_params": [_P],
        "qb_set_optimizer": [_P, _f64, _f64],
        "qb_set_data": [_P, _P, _i32, _P, _i32, _i64],
        "qb_init_chains": [_P, _i64, _P, _P, _P],
        "qb_get_chains": [_P, _P, _P, _P],
        "qb_get_chain_rows": [_P, _i64, _i64, _P, _P, _P],
        "qb_get_gemm_histogram": [_P, _P, _i32],
        "qb_get_issued_flops": [_P, _P],
        "qb_seed": [_P, ctypes.c_uint64],
        "qb_inject_streams": [_P, _P, _P, _P, _P, _i64, _i64],
        "qb_pcd_step": [_P, _i64, _i64, _i64, _f64, _f64],
        "qb_cd_step": [_P, _i64, _i64, _i64, _f64, _f64],
        "qb_pcd_epoch": [_P, _i64, _i64, _f64, _f64, _P],
        "qb_cd_epoch": [_P, _i64, _i64, _f64, _f64, _P],
        "qb_train_steps": [_P, _i32, _i64, _i64, _i64, _i64, _f64, _f64],
        "qb_fast_pcd_epoch": [_P, _i64, _i64, _f64, _f64, _f64, _f64, _P],
        "qb_pcd_step_host": [_P, _P, _i32, _P, _i32, _i64, _i64, _f64, _f64, _P],
        "qb_cd_step_host": [_P, _P, _i32, _P, _i32, _i64, _i64, _f64, _f64, _P],
        "qb_set_next_host_batch": [_P, _P, _i32, _P, _i32],
        "qb_eval_mse": [_P, _P, _i32, _i64, _P, _P],
        "qb_eval_free_energy": [_P, _P, _i32, _P, _i32, _i64, _P],
        "qb_reconstruct": [_P, _P, _i64, _P, _P],
        "qb_conditional_prob_h": [_P, _P, _P, _i64, _P],
        "qb_sample_hidden": [_P, _P, _P, _i64, _P, _P],
        "qb_sample_visible": [_P, _P, _i64, _P, _P, _P, _P],
        "qb_conditional_prob_y_given_v": [_P, _P, _i32, _i64, _P],
        "qb_comm_unique_id": [_P],
        "qb_comm_init": [_P, _P],
        "qb_p2p_export": [_P, _P],
        "qb_p2p_import": [_P, _P],
        "qb_get_p2p_timing": [_P, _P, _P],
        "qb_get_counters": [_P, _P, _P, _P, _P, _P, _P],
        "qb_get_update_timing": [_P, _P, _P],
        "qb_sync": [_P],
        "qb_timer_start": [_P],
        "qb_timer_stop": [_P, _P],
        "qb_reset_counters": [_P],
        "qb_train": [_P, ctypes.POINTER(QbTrainConfig), _P, _P, _P],
        "qb_set_option": [_P, ctypes.c_char_p, _f64],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = _i32
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise QbError(rc, lib().qb_last_error().decode("utf-8", "replace"))


_DT = {np.dtype(np.float64): QB_DT_F64, np.dtype(np.int64): QB_DT_I64, np.dtype(np.float32): QB_DT_F32,
       np.dtype(np.uint8): QB_DT_U8}


def as_rows(x, cols: int):
    """Vector{Vector{T}} / 2-D array -> (C-contiguous [n][cols] ndarray, QB_DT code)."""
    if x is None:
        return None, 0
    if hasattr(x, "data_ptr"):  # torch CPU tensor (possibly pinned)
        x = x.numpy()
    a = np.asarray(x)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype not in _DT:
        a = a.astype(np.int64) if np.issubdtype(a.dtype, np.integer) else a.astype(np.float64)
    a = np.ascontiguousarray(a).reshape(-1, cols)
    return a, _DT[a.dtype]


def ptr(a):
    return None if a is None else a.ctypes.data_as(_P)
