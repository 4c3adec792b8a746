"""Reference-exact online CD (batch size 1, train_cd.jl:2-21) at the profiling/mnist.jl shape: one cooperative
kernel per epoch vs one launch sequence per sample."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qarbom_b200 as q

V, H = 784, 500
rng = np.random.default_rng(0)
for k in (1, 3):
    for online, n, wpb in ((1, 20000, 4), (1, 20000, 8), (1, 20000, 16), (1, 20000, 32), (0, 2000, 0)):
        X = (rng.random((n, V)) < 0.1307).astype(np.uint8)
        rbm = q.RBM(V, H, W=0.01 * rng.standard_normal((V, H)))
        with q.DeviceRBM(rbm, max_batch=4, precision="bf16", seed=1) as dev:
            dev.set_option("online_kernel", online)
            dev.set_option("online_wpb", wpb)
            dev.set_data(X)
            dev.cd_epoch(1, k, 0.01, timed=False)
            t0 = time.perf_counter()
            t = dev.cd_epoch(1, k, 0.01, timed=bool(online))
            dt = time.perf_counter() - t0
            print(f"k={k} online_kernel={online} wpb={wpb}: {n / dt:10.0f} samples/s  ({dt / n * 1e6:6.2f} us/sample)  phases(s)={tuple(round(x, 4) for x in t)}", flush=True)
