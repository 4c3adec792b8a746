"""Sweep the GEMM tile shape (force_bn x use_2cta) on the small BASELINE configs: ms per step for each choice
against the cost-model default.  Usage: python tools/shape_sweep.py [cfg3 cfg4 ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qarbom_b200 as q

CFG = {  # V, H, L, gaussian, B, k, pcd
    "cfg1": (784, 500, 0, False, 100, 1, False),
    "cfg2r": (784, 500, 0, False, 256, 1, True),
    "cfg3": (1024, 2048, 0, True, 1024, 5, False),
    "cfg4": (784, 1000, 10, False, 512, 1, False),
}
names = sys.argv[1:] or ["cfg3", "cfg4"]
for name in names:
    V, H, L, gaussian, B, k, pcd = CFG[name]
    rng = np.random.default_rng(0)
    nb = 4
    X = rng.standard_normal((nb * B, V)).astype(np.float32) if gaussian else (rng.random((nb * B, V)) < 0.13).astype(np.uint8)
    Y = np.eye(L, dtype=np.uint8)[np.arange(nb * B) % L] if L else None
    W = 0.01 * rng.standard_normal((V, H))
    if L:
        rbm = (q.GRBMClassifier if gaussian else q.RBMClassifier)(V, H, L, W=W, U=0.01 * rng.standard_normal((L, H)))
    else:
        rbm = (q.GRBM if gaussian else q.RBM)(V, H, W=W)
    dev = q.DeviceRBM(rbm, max_batch=B, precision="bf16", seed=1)
    dev.set_data(X, Y)
    dev.init_chains(B)
    dev.set_option("async_steps", 1)
    step = (lambda i: dev.pcd_step(i % nb, B, k, 0.01, 0.01)) if pcd else (lambda i: dev.cd_step(i % nb, B, k, 0.01, 0.01))
    combos = [(0, 1)] + [(bn, two) for two in (0, 2) for bn in (32, 64, 128, 256) if not (two == 2 and bn < 128)]
    for bn, two in combos:
        dev.set_option("force_bn", bn); dev.set_option("use_2cta", two)
        try:
            for i in range(5):
                step(i)
            dev.sync()
            best = 1e9
            for rep in range(3):
                dev.timer_start()
                n = 200
                for i in range(n):
                    step(i)
                best = min(best, dev.timer_stop() / n)
            print(f"{name} force_bn={bn:3d} use_2cta={two}  {best * 1e3:8.1f} us/step", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{name} force_bn={bn} use_2cta={two}  failed: {e}", flush=True)
    dev.close()
