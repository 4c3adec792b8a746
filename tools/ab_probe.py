"""A/B in the sustained (power-capped) regime: alternate kernel options, 200 cfg5 steps each."""
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
configs = [("1cta", dict(use_2cta=0)), ("2cta", dict(use_2cta=1))]
extra = sys.argv[1:]
for rep in range(3):
    for name, opts in configs:
        for kk, vv in opts.items():
            dev.set_option(kk, vv)
        for i in range(3):
            dev.pcd_step(i % nb, B, k, 0.01)
        dev.sync()
        dev.timer_start()
        n = 200
        for i in range(n):
            dev.pcd_step(i % nb, B, k, 0.01)
        ms = dev.timer_stop()
        print(f"rep {rep} {name:8s} {ms / n:7.3f} ms/step  {B * n / ms / 1e3:7.3f} M samples/s", flush=True)
