"""CPU-side checks of the boundary: the library builds, loads, exports every symbol that
include/qarbom_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def q():
    import qarbom_b200
    qarbom_b200.build()
    return qarbom_b200


def test_header_symbols_are_exported(q):
    hdr = open(os.path.join(ROOT, "include", "qarbom_b200.h")).read()
    declared = set(re.findall(r"\b(qb_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"qb_handle", "qb_config"}
    assert declared == set(q._ffi.SYMBOLS), declared ^ set(q._ffi.SYMBOLS)
    L = q.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.qb_version() == 100


def test_no_cpu_fallback(q):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(q.QbError) as e:
        q.DeviceRBM(q.RBM(3, 2), max_batch=2)
    assert e.value.code == -2  # QB_E_CUDA


def test_host_model_types_mirror_reference():
    import qarbom_b200 as q
    r = q.RBM(5, 3, rng=np.random.default_rng(0))
    assert r.W.shape == (5, 3) and not r.a.any() and not r.b.any() and abs(r.W.std() - 1.0) < 0.6  # randn, sigma = 1
    c = q.RBMClassifier(5, 3, 2, rng=np.random.default_rng(0))
    assert c.U.shape == (2, 3) and c.n_classifiers == 2
    g = q.GRBMClassifier(5, 3, 2, W=np.zeros((5, 3)), U=np.zeros((2, 3)))
    assert np.all(g.a == 1e-5) and np.all(g.c == 1e-5)  # src/rbm.jl:94-99
    with pytest.raises(TypeError):
        q.train(r, [[0] * 5] * 4, q.PCD, n_epochs=1, learning_rate=[0.1])  # batch_size has no default


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "qarbom.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f


def test_bench_roofline_traffic_reads_the_committed_ncu_summaries():
    """bench.py's roofline.traffic comes from the `ncu --set full` summaries committed under profiles/ (tools/ncu_summary.py):
    the files it names must exist and parse to a per-launch byte count for the kernels of the timed region."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name in (bench.NCU_GEMM_SUMMARY, bench.NCU_UPDATE_SUMMARY):
        assert os.path.exists(os.path.join(root, "profiles", name)), name
    gemm = bench.ncu_traffic(bench.NCU_GEMM_SUMMARY, "gibbs_", mix=bench.gemm_mix(bench.WORKLOADS["cfg5"]))
    upd = bench.ncu_traffic(bench.NCU_UPDATE_SUMMARY, "update_w_kernel")
    # algorithmic bytes: a chain launch streams >= 21 x 2 x 67 MB of states, the update 268 MB
    assert gemm is not None and 1e9 < gemm < 2e10, gemm
    assert upd is not None and 1e8 < upd < 1e9, upd
