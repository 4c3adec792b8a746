"""What one rank of an N-GPU cfg5 run does, on ONE GPU without NCCL: B_local = 8192 / N rows."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qarbom_b200 as q

V = H = 4096; k = 10; nb = 2
rng = np.random.default_rng(0)
rbm = q.RBM(V, H, W=0.01 * rng.standard_normal((V, H)))
for B in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]:
    X = (rng.random((nb * B, V), dtype=np.float32) < 0.5).astype(np.uint8)
    dev = q.DeviceRBM(rbm, max_batch=B, precision="bf16", seed=1)
    dev.set_data(X); dev.init_chains(B)
    dev.set_option("async_steps", 1)
    for i in range(5):
        dev.pcd_step(i % nb, B, k, 0.01)
    dev.sync()
    n = 100
    t0 = time.perf_counter()
    dev.timer_start()
    for i in range(n):
        dev.pcd_step(i % nb, B, k, 0.01)
    t_host = time.perf_counter() - t0
    ms = dev.timer_stop()
    dev.reset_counters(); dev.set_option("time_gemms", 1)
    for i in range(20):
        dev.pcd_step(i % nb, B, k, 0.01)
    c = dev.counters_full()
    print(f"B={B:5d}  {ms / n:7.3f} ms/step (host enqueue {t_host / n * 1e3:6.3f} ms/step)   avg gemm {c[4] / c[3] * 1e3:7.1f} us  "
          f"({c[5] / 1e12 / (c[4] * 1e-3):7.1f} TF/s)", flush=True)
    dev.close()
