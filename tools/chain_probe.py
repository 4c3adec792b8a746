"""Fused chain kernel vs one launch per half-step, and chain tile widths, on ONE GPU.
usage: chain_probe.py <workload> [rows ...]     workload: cfg5 (PCD k=10, 4096x4096) | cfg3 (GRBM CD-5 1024x2048) | cfg1 | cfg4"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qarbom_b200 as q

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
CFG = {"cfg5": dict(V=4096, H=4096, L=0, g=False, k=10, pcd=True, rows=[1024, 2048, 8192]),
       "cfg3": dict(V=1024, H=2048, L=0, g=True, k=5, pcd=False, rows=[1024]),
       "cfg1": dict(V=784, H=500, L=0, g=False, k=1, pcd=False, rows=[100]),
       "cfg2": dict(V=784, H=500, L=0, g=False, k=1, pcd=True, rows=[256]),
       "cfg4": dict(V=784, H=1000, L=10, g=False, k=1, pcd=False, rows=[512])}[wl]
V, H, L, k = CFG["V"], CFG["H"], CFG["L"], CFG["k"]
rng = np.random.default_rng(0)
W = 0.01 * rng.standard_normal((V, H))
if L:
    rbm = q.RBMClassifier(V, H, L, W=W, U=0.01 * rng.standard_normal((L, H)))
else:
    rbm = (q.GRBM if CFG["g"] else q.RBM)(V, H, W=W)
nb = 4
for B in [int(a) for a in sys.argv[2:]] or CFG["rows"]:
    X = rng.standard_normal((nb * B, V)).astype(np.float32) if CFG["g"] else (rng.random((nb * B, V), dtype=np.float32) < 0.5).astype(np.uint8)
    Y = np.eye(L, dtype=np.uint8)[np.arange(nb * B) % L] if L else None
    modes = [(0, 0, 0), (1, 0, 0), (2, 128, 0), (2, 256, 0), (2, 0, 2)] if os.environ.get("CHAIN_PROBE_SHORT") else \
            [(0, 0, 0), (1, 0, 0), (1, 0, 1), (2, 64, 0), (2, 128, 0), (2, 256, 0), (2, 0, 2)]
    pairs_list = [int(x) for x in os.environ.get("CHAIN_PAIRS", "0").split(",")]   # option "chain_pairs": 0 = all CTA pairs
    modes = [(f, b, qd, p) for (f, b, qd) in modes for p in (pairs_list if f else [0])]
    for fused, bn, quad, pairs in modes:   # quad: 4-CTA-cluster chain kernel with multicast weights (0 never, 1 automatic, 2 whenever possible)
        if quad == 2 and B % 256:
            continue
        dev = q.DeviceRBM(rbm, max_batch=B, precision="bf16", seed=1)
        dev.set_option("fused_chain", fused); dev.set_option("chain_bn", bn); dev.set_option("chain_quad", quad); dev.set_option("chain_pairs", pairs)
        dev.set_data(X, Y)
        if CFG["pcd"]:
            dev.init_chains(B)
        step = (lambda i: dev.pcd_step(i % nb, B, k, 0.01, 0.01)) if CFG["pcd"] else (lambda i: dev.cd_step(i % nb, B, k, 0.01, 0.01))
        dev.set_option("async_steps", 1)
        for i in range(5):
            step(i)
        dev.sync()
        n = max(20, min(2000, int(0.3 / (3e-5 + 5e-10 * B * (V + L) * H * (2 * k + 3) / 1e3))))
        dev.reset_counters()
        dev.timer_start()
        for i in range(n):
            step(i)
        ms = dev.timer_stop()
        launches = dev.counters()[0] / n
        print(f"{wl} B={B:5d} fused={fused} chain_bn={bn:3d} quad={quad} pairs={pairs:3d}: {ms / n * 1e3:9.1f} us/step  {launches:4.1f} launches/step  hist={dev.gemm_histogram()}", flush=True)
        dev.close()
