"""Persistent multi-mini-batch kernel vs one launch sequence per step on the small BASELINE configs (ONE GPU).
usage: epoch_probe.py [cfg1 cfg2 cfg3 cfg4]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qarbom_b200 as q

CFG = {"cfg1": dict(V=784, H=500, L=0, g=False, k=1, algo="CD", B=100, nb=600),
       "cfg2": dict(V=784, H=500, L=0, g=False, k=1, algo="PCD", B=256, nb=234),
       "cfg3": dict(V=1024, H=2048, L=0, g=True, k=5, algo="CD", B=1024, nb=64),
       "cfg4": dict(V=784, H=1000, L=10, g=False, k=1, algo="CD", B=512, nb=117)}
for wl in sys.argv[1:] or list(CFG):
    c = CFG[wl]
    V, H, L, k, B, nb = c["V"], c["H"], c["L"], c["k"], c["B"], c["nb"]
    rng = np.random.default_rng(0)
    W = 0.01 * rng.standard_normal((V, H))
    rbm = q.RBMClassifier(V, H, L, W=W, U=0.01 * rng.standard_normal((L, H))) if L else (q.GRBM if c["g"] else q.RBM)(V, H, W=W)
    X = rng.standard_normal((nb * B, V)).astype(np.float32) if c["g"] else (rng.random((nb * B, V), dtype=np.float32) < 0.13).astype(np.uint8)
    Y = np.eye(L, dtype=np.uint8)[np.arange(nb * B) % L] if L else None
    for mode, bn in [(0, 0), (2, 32), (2, 64)]:
        dev = q.DeviceRBM(rbm, max_batch=B, precision="bf16", seed=1)
        dev.set_option("epoch_kernel", mode)
        if bn:
            dev.set_option("force_bn", bn)
        dev.set_data(X, Y)
        if c["algo"] == "PCD":
            dev.init_chains(B)
        dev.set_option("async_steps", 1)
        dev.train_steps(c["algo"], 0, min(nb, 50), B, k, 0.01, 0.01)
        dev.sync()
        n = 2 * nb
        dev.reset_counters()
        dev.timer_start()
        dev.train_steps(c["algo"], 0, n, B, k, 0.01, 0.01)
        ms = dev.timer_stop()
        print(f"{wl} B={B} epoch_kernel={mode} bn={bn:2d}: {ms / n * 1e3:8.2f} us/step  {dev.counters()[0] / n:5.2f} launches/step  hist={dev.gemm_histogram()}", flush=True)
        dev.close()
