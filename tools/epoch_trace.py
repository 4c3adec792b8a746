"""Per-stage timeline of the persistent multi-mini-batch kernel (development aid, ONE GPU).

Needs a library built with -DQB_EPOCH_TRACE (python tools/epoch_trace.py --build writes lib/libqarbom_b200_trace.so; the
shipped library has no stamps).  usage: epoch_trace.py [--build] [cfg1|cfg2|cfg4]

Each tile writes %globaltimer at: 0 producer picks the tile up, 1 whole-step waits done, 2 row-block wait done,
3 first operands landed (MMA starts), 4 last operands landed, 5 epilogue starts, 6 epilogue math done, 7 counters released,
8+kb (kb < 20) SM clock words of k-step kb: operands landed (low), MMAs issued (high); 28 clocks spent issuing the state loads.
The report is, per step of the iteration, the median over iterations of (first start, last release) relative to the
iteration's first event, and the mean width of each phase on the tile that finishes last."""
import os, struct, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "qarbom.jl_b200", "lib", "libqarbom_b200_trace.so")

if "--build" in sys.argv:
    sys.path.insert(0, os.path.join(ROOT, "qarbom.jl_b200"))
    import build
    src = os.path.join(ROOT, "qarbom.jl_b200", "csrc")
    subprocess.check_call(["nvcc", *build.NVCC_FLAGS, *build._nccl_device_api_flags(), "-DQB_EPOCH_TRACE", os.path.join(src, "qb_api.cu"),
                           os.path.join(src, "nvls_setup.cu"), "-o", LIB, "-lcudart", "-ldl"])
    print("built", LIB)
    sys.exit(0)

os.environ["QARBOM_B200_LIB"] = LIB
import qarbom_b200 as q

CFG = {"cfg1": dict(V=784, H=500, L=0, k=1, algo="CD", B=100, nb=64),
       "cfg2": dict(V=784, H=500, L=0, k=1, algo="PCD", B=256, nb=64),
       "cfg4": dict(V=784, H=1000, L=10, k=1, algo="CD", B=512, nb=64)}
wl = next((a for a in sys.argv[1:] if a in CFG), "cfg1")
c = CFG[wl]
V, H, L, k, B, nb = c["V"], c["H"], c["L"], c["k"], c["B"], c["nb"]
rng = np.random.default_rng(0)
W = 0.01 * rng.standard_normal((V, H))
rbm = q.RBMClassifier(V, H, L, W=W, U=0.01 * rng.standard_normal((L, H))) if L else q.RBM(V, H, W=W)
X = (rng.random((nb * B, V), dtype=np.float32) < 0.13).astype(np.uint8)
Y = np.eye(L, dtype=np.uint8)[np.arange(nb * B) % L] if L else None
dev = q.DeviceRBM(rbm, max_batch=B, precision="bf16", seed=1)
dev.set_option("epoch_kernel", 2)
for a in sys.argv[1:]:
    if a.startswith("bn="):
        dev.set_option("force_bn", int(a[3:]))
dev.set_data(X, Y)
if c["algo"] == "PCD":
    dev.init_chains(B)
dev.train_steps(c["algo"], 0, 32, B, k, 0.01, 0.01)
dev.sync()
path = os.path.join(ROOT, "gpurun_out", f"epoch_trace_{wl}.bin")
os.makedirs(os.path.dirname(path), exist_ok=True)
os.environ["QB_EPOCH_TRACE"] = path
n = 128
dev.timer_start()
dev.train_steps(c["algo"], 0, n, B, k, 0.01, 0.01)
ms = dev.timer_stop()
dev.close()
raw = open(path, "rb").read()
hdr = struct.unpack("<30q", raw[:240])
n_iters, tpi, n_steps = hdr[0], hdr[1], hdr[2]
base = list(hdr[4:4 + n_steps]) + [tpi]
t_raw = np.frombuffer(raw[240:], dtype=np.uint64).reshape(n_iters, tpi, 32)
t = t_raw.astype(np.float64)
t[t == 0] = np.nan
print(f"{wl}: {n_iters} iterations, {tpi} tiles per iteration in {n_steps} steps (tile bases {base[:-1]}), {ms / n * 1e3:.2f} us per iteration with stamps")
it_len = np.diff(np.nanmax(t[:, :, 7], axis=1))
print(f"iteration period (release of the last tile): median {np.median(it_len) / 1e3:.2f} us")
names = ["pick", "dep", "rowblk", "mma0", "mmaN", "epi0", "epi1", "rel"]
for s in range(n_steps):
    seg = t[8:, base[s]:base[s + 1], :]          # skip the ramp
    t0 = np.nanmin(t[8:, :, 1], axis=1)[:, None]  # first tile through its whole-step wait = the iteration's start
    first = np.nanmin(np.where(np.isnan(seg[:, :, 2]), seg[:, :, 1], seg[:, :, 2]), axis=1) - t0[:, 0]
    last = np.nanmax(seg[:, :, 7], axis=1) - t0[:, 0]
    crit = np.nanargmax(seg[:, :, 7], axis=1)
    rows = seg[np.arange(seg.shape[0]), crit, :]
    ready = np.where(np.isnan(rows[:, 2]), rows[:, 1], rows[:, 2])
    w = [np.nanmedian(rows[:, 3] - ready), np.nanmedian(rows[:, 4] - rows[:, 3]), np.nanmedian(rows[:, 5] - rows[:, 4]),
         np.nanmedian(rows[:, 6] - rows[:, 5]), np.nanmedian(rows[:, 7] - rows[:, 6])]
    print(f"step {s}: tiles {base[s + 1] - base[s]:3d}  first ready {np.median(first) / 1e3:6.2f} us  last release {np.median(last) / 1e3:6.2f} us | last tile: "
          f"ready->operands {w[0] / 1e3:5.2f}  k-loop loads {w[1] / 1e3:5.2f}  ->epilogue {w[2] / 1e3:5.2f}  epilogue {w[3] / 1e3:5.2f}  fence+release {w[4] / 1e3:5.2f}")
    raw_kb = t_raw[8:, base[s]:base[s + 1], 8:28][np.arange(seg.shape[0]), crit, :]
    landed = (raw_kb & np.uint64(0xFFFFFFFF)).astype(np.int64); issued = (raw_kb >> np.uint64(32)).astype(np.int64)
    nk = int((raw_kb[0] != 0).sum())
    wait = np.median((landed[:, 1:nk] - issued[:, :nk - 1]) & 0xFFFFFFFF, axis=0)
    issue = np.median((issued[:, :nk] - landed[:, :nk]) & 0xFFFFFFFF, axis=0)
    print("        SM clocks per k-step, issuing 4 tcgen05.mma + commit: " + " ".join(f"{a:.0f}" for a in issue))
    print("        SM clocks from there to the next k-step's operands:   " + " ".join(f"{a:.0f}" for a in wait) + f"   | producer, issuing the state loads after the wait: {np.nanmedian(rows[:, 28]):.0f}")
