"""Experiment driver: cfg5-shaped PCD steps under different kernel options (prints ms/step and avg GEMM ms)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qarbom_b200 as q

V = H = 4096; B = 8192; k = 10; nb = 2
rng = np.random.default_rng(0)
X = (rng.random((nb * B, V), dtype=np.float32) < 0.5).astype(np.uint8)
rbm = q.RBM(V, H, W=0.01 * rng.standard_normal((V, H)))
dev = q.DeviceRBM(rbm, max_batch=B, precision="bf16", seed=1)
dev.set_data(X); dev.init_chains(B)
dev.set_option("async_steps", 1)
for name, opts in [("1cta", dict(use_2cta=0)), ("2cta", dict(use_2cta=1)),
                   ("1cta nostore", dict(use_2cta=0, dbg=1)), ("1cta nophilox", dict(use_2cta=0, dbg=2)),
                   ("1cta nosigmoid", dict(use_2cta=0, dbg=4)), ("1cta none", dict(use_2cta=0, dbg=7)),
                   ("2cta nostore", dict(use_2cta=1, dbg=1)), ("2cta nophilox", dict(use_2cta=1, dbg=2)), ("2cta none", dict(use_2cta=1, dbg=7))]:
    dev.set_option("dbg", 0)
    for kk, vv in opts.items():
        dev.set_option(kk, vv)
    for i in range(2):
        dev.pcd_step(i % nb, B, k, 0.01)
    dev.sync(); dev.reset_counters()
    dev.timer_start()
    n = 6
    for i in range(n):
        dev.pcd_step(i % nb, B, k, 0.01)
    ms = dev.timer_stop()
    dev.reset_counters(); dev.set_option("time_gemms", 1)
    for i in range(3):
        dev.pcd_step(i % nb, B, k, 0.01)
    c = dev.counters_full()
    dev.set_option("time_gemms", 0)
    print(f"{name:16s} {ms / n:7.3f} ms/step   avg gemm {c[4] / c[3] * 1e3:7.1f} us  ({c[5] / 1e12 / (c[4] * 1e-3):7.1f} TF/s)")
