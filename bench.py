#!/usr/bin/env python
"""Benchmark of the RBM training hot path (CD-k / PCD) on B200, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5] [--impl ours|reference]

A "step" is one mini-batch step of the hot path (k Gibbs steps on the persistent chains,
positive phase, gradient, allreduce, update).  The default workload is BASELINE.json's
configs[4] -- Binary RBM 4096x4096, PCD k=10, global batch 8192 -- the configuration the
headline metric ("... at 1/2/4/8 B200; % of roofline") and the north-star targets are
quoted on; it fits one GPU.  With N > 1 the global batch is sharded by rows/chains over the
ranks (strong scaling, as the config says) with one NCCL gradient allreduce per step.

`value`   : samples/s with the dataset already resident in HBM (device-timed, CUDA events on
            the library's stream, max over ranks).
`e2e`     : the same metric through qb_pcd_step_host: every step uploads its mini-batch from
            pinned host memory and reads the batch reconstruction error back.
`roofline`: algorithmic GEMM flops / CUDA-event time of the GEMM launches inside the timed
            region, against MEASURED_PEAKS.json (sustained bf16: the kernel runs inside a long step).
`cpu_baseline` / `--impl reference`: the C restatement of the reference's per-sample Float64
            algorithm (oracle/, Julia is not installable here) on the host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: V, H, L, gaussian, global batch, k, algo, n_batches, data density
    "cfg1": dict(V=784, H=500, L=0, gaussian=False, B=100, k=1, algo="CD", nb=600, density=0.1307,
                 desc="Binary RBM 784x500, CD-1, batch 100"),
    "cfg2": dict(V=784, H=500, L=0, gaussian=False, B=256, k=1, algo="PCD", nb=234, density=0.1307, chains=1000,
                 desc="Binary RBM 784x500, PCD with 1000 persistent chains, k=1, batch 256"),
    "cfg2r": dict(V=784, H=500, L=0, gaussian=False, B=256, k=1, algo="PCD", nb=234, density=0.1307,
                  desc="Binary RBM 784x500, PCD k=1, batch 256, chains = batch (the reference's own semantics)"),
    "cfg3": dict(V=1024, H=2048, L=0, gaussian=True, B=1024, k=5, algo="CD", nb=64, density=None,
                 desc="Gaussian-Bernoulli RBM 1024x2048, CD-5, batch 1024"),
    "cfg4": dict(V=784, H=1000, L=10, gaussian=False, B=512, k=1, algo="CD", nb=117, density=0.1307,
                 desc="RBMClassifier 784+10 x 1000, CD-1, batch 512"),
    "cfg5": dict(V=4096, H=4096, L=0, gaussian=False, B=8192, k=10, algo="PCD", nb=8, density=0.5,
                 desc="Binary RBM 4096x4096, PCD k=10, batch 8192"),
}
METRIC = "PCD/CD-k training samples/sec"
# `ncu --set full` summaries of this round (tools/ncu_summary.py raw <rep>), committed under profiles/
NCU_GEMM_SUMMARY = "r2b_ncu_gemm_raw_summary.txt"
NCU_UPDATE_SUMMARY = "r2b_ncu_update_raw_summary.txt"


def gemm_mix(w):
    """Launches per step by kernel specialisation (template argument text of the ncu kernel name)."""
    # With the persistent chain kernel (the default where a half-step is at least one wave of pair tiles, e.g. cfg5) a step is
    # ONE gibbs_chain_kernel launch (all 2k + 1 forward contractions) + the gradient GEMM; the per-half-step kernels appear in a
    # capture only when that kernel is off.  Keys absent from the capture drop out of the weighted mean.
    k = w["k"]
    if w["algo"] == "PCD":
        return {"gibbs_chain_kernel": 1, ", 2>": 2 * k, ", 3>": 1, ", 1>": 1}      # chain | K_FAST_SAMPLE, K_FAST_PROB, K_GRAD
    return {"gibbs_chain_kernel": 1, ", 2>": 2 * k - 1, ", 4>": 1, ", 3>": 1, ", 1>": 1}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tf=float(d["bf16_tflops_sustained"]), tf_burst=float(d["bf16_tflops"]), hbm=float(d["hbm_gbs"]),
                    src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tf=1400.0, tf_burst=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        after = False
        if not sm:  # timed region shorter than nvidia-smi's start-up: one query right after it
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
                r = [x.strip() for x in out.split(",")]
                sm.append(float(r[0])); mx = float(r[1]); after = True
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "sampled_after_region": after}


def ncu_traffic(fname: str, kernel_prefix: str, mix=None):
    """dram read+write bytes per launch from a committed `ncu --set full` summary (profiles/, tools/ncu_summary.py).
    `mix` = {template-argument substring: launches per step} weights the per-specialisation means by the launch mix of the
    step (e.g. 20 sampling : 1 probability : 1 gradient GEMM); without it, the plain mean over the captured launches.
    None if no capture is committed."""
    path = os.path.join(ROOT, "profiles", fname)
    if not os.path.exists(path):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}   # kernel name -> [bytes, launches]
    cur = None
    for line in open(path):
        if not line.startswith(" "):
            cur = line.strip() if kernel_prefix in line else None
            if cur:
                per.setdefault(cur, [0.0, 0])[1] += 1
        elif cur and ("dram__bytes_read.sum " in line or "dram__bytes_write.sum " in line):
            parts = line.split()
            per[cur][0] += float(parts[1]) * scale.get(parts[2], 1.0)
    if not per:
        return None
    if not mix:
        return sum(b for b, _ in per.values()) / sum(n for _, n in per.values())
    tot, wsum = 0.0, 0.0
    for key, weight in mix.items():
        hit = [(b, n) for name, (b, n) in per.items() if key in name and n]
        if hit:
            tot += weight * sum(b for b, _ in hit) / sum(n for _, n in hit)
            wsum += weight
    return tot / wsum if wsum else None


def host_cores() -> int:
    """Cores this process may run on (NOT omp_get_max_threads: torch.distributed.run exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def make_data(w, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    n = w["B"] * w["nb"]
    if w["gaussian"]:
        X = rng.standard_normal((n, w["V"]), dtype=np.float32)
    else:
        X = (rng.random((n, w["V"]), dtype=np.float32) < w["density"]).astype(np.uint8)
    Y = None
    if w["L"]:
        Y = np.zeros((n, w["L"]), dtype=np.uint8)
        Y[np.arange(n), np.arange(n) % w["L"]] = 1
    W0 = 0.01 * rng.standard_normal((w["V"], w["H"]))
    U0 = 0.01 * rng.standard_normal((w["L"], w["H"])) if w["L"] else None
    return X, Y, W0, U0


def shard_rows(A, B, nb, rank, world):
    """Rank's contiguous slice of every global mini-batch, concatenated in mini-batch order."""
    if A is None:
        return None
    Bl = B // world
    A = A.reshape(nb, B, -1)[:, rank * Bl:(rank + 1) * Bl]
    return np.ascontiguousarray(A.reshape(nb * Bl, -1))


def flops_per_step(w):
    """(2k+3) * 2 * B * (V+L) * H (de-duplicated forward + 2 gradient contractions);
    with Nc != B persistent chains the 2k chain half-steps and one gradient product run over Nc rows."""
    vh = 2.0 * (w["V"] + w["L"]) * w["H"]
    nc = w.get("chains", w["B"])
    return (2 * w["k"] + 1) * nc * vh + 2 * w["B"] * vh if w["algo"] == "PCD" else (2 * w["k"] + 3) * w["B"] * vh


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference(w, seconds_target, steps, warmup, threads=None):
    """Times the C restatement of the reference on a bounded sample: `rows` rows of one
    mini-batch per step (per-sample cost of the reference does not depend on the batch size)."""
    from oracle import c_oracle
    threads = threads or host_cores()   # explicit num_threads() in the C code: independent of torchrun's OMP_NUM_THREADS=1
    V, H, L, k = w["V"], w["H"], w["L"], w["k"]
    rng = np.random.Generator(np.random.PCG64(7))
    W0 = 0.01 * rng.standard_normal((V, H))
    U0 = 0.01 * rng.standard_normal((L, H)) if L else None

    def run(rows, n_steps):
        m = c_oracle.CModel(W0, np.zeros(V), np.zeros(H), U0, np.zeros(L) if L else None, w["gaussian"])
        X = rng.standard_normal((rows, V)) if w["gaussian"] else (rng.random((rows, V)) < w["density"]).astype(np.float64)
        Y = np.eye(L)[np.arange(rows) % L] if L else None
        nch = max(1, int(round(rows * w.get("chains", w["B"]) / w["B"])))  # keep the chains : rows ratio of the workload
        chv, chh = rng.random((nch, V)), rng.random((nch, H))
        chy = rng.random((nch, L)) if L else None
        ts = []
        for _ in range(n_steps):
            t0 = time.perf_counter()
            if w["algo"] == "PCD":
                c_oracle.pcd_epoch(m, X, Y, rows, k, 0.01, 0.01, chv, chh, chy, seed=1, threads=threads, n_chains=nch)
            else:
                c_oracle.cd_epoch(m, X, Y, rows, k, 0.01, 0.01, seed=1, threads=threads)
            ts.append(time.perf_counter() - t0)
        return ts

    t_probe = run(2, 2)[1] / 2  # seconds per row (second call: thread pool and pages are warm)
    per_step = max(seconds_target / max(steps + warmup, 1), 0.05)
    rows = int(max(2, min(w["B"], per_step / max(t_probe, 1e-9))))
    ts = run(rows, warmup + steps)[warmup:]
    sps = rows * len(ts) / sum(ts)
    return dict(value=sps, rows=rows, threads=threads, ms_per_step=1e3 * sum(ts) / len(ts),
                sample=f"{rows} rows of one {w['B']}-row mini-batch per step (k={k}), {len(ts)} timed steps; "
                       f"per-sample cost of the reference is independent of the batch size")


def cpu_batched_blas(w, seconds_target):
    """Best-case CPU figure: the same step in the batched BLAS-3 Float64 formulation (oracle `batched_step`, numpy's
    threaded dgemm), on a bounded sample of rows.  Not the reference's algorithm shape (it is per-sample gemv), reported
    next to it so that the GPU/CPU ratio can be read against a CPU that is used well."""
    from oracle import qarbom_oracle as qo
    threads = host_cores()
    try:   # numpy's BLAS pool may have been sized by OMP_NUM_THREADS=1 (torchrun): resize it to the host's cores
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=threads)
        threads = max([d.get("num_threads", 1) for d in threadpool_info()] or [threads])
    except Exception:  # noqa: BLE001
        pass
    V, H, L, k = w["V"], w["H"], w["L"], w["k"]
    rng = np.random.default_rng(7)
    rows = int(min(w["B"], 1024))
    nch = max(1, int(round(rows * w.get("chains", w["B"]) / w["B"])))
    m = qo.make_model(V, H, L, w["gaussian"], W=0.01 * rng.standard_normal((V, H)), U=0.01 * rng.standard_normal((L, H)) if L else None)
    X = rng.standard_normal((rows, V)) if w["gaussian"] else (rng.random((rows, V)) < w["density"]).astype(np.float64)
    Y = np.eye(L)[np.arange(rows) % L] if L else None
    cv, ch = rng.random((nch, V)), rng.random((nch, H))
    cy = rng.random((nch, L)) if L else None

    def step():
        nonlocal cv, ch, cy
        with np.errstate(all="ignore"):
            if w["algo"] == "PCD" and nch != rows:   # more chains than rows: the extra chains only advance (cost-equivalent stand-in)
                cv, ch, cy = _more(cv, ch, cy)
            else:
                cv, ch, cy = qo.batched_step(m, X, Y, w["algo"], k, 0.01, 0.01, cv=cv, ch=ch, cy=cy, rng=rng)

    def _more(cv, ch, cy):
        extra_v, extra_h = cv[rows:], ch[rows:]
        for _ in range(k):   # the chains beyond the mini-batch rows only advance
            extra_h = (rng.random(extra_h.shape) < qo.logistic(m.b + extra_v @ m.W)).astype(np.float64)
            extra_v = (rng.random(extra_v.shape) < qo.logistic(m.a + extra_h @ m.W.T)).astype(np.float64)
        a, b, c = qo.batched_step(m, X, Y, "PCD", k, 0.01, 0.01, cv=cv[:rows], ch=ch[:rows], cy=None, rng=rng)
        return np.concatenate([a, extra_v]), np.concatenate([b, extra_h]), c

    step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds_target or n >= 200:
            break
    return dict(value=rows * n / dt, threads=threads,
                sample=f"{rows} rows per step (k={k}), {n} timed steps, numpy Float64 matmul on {threads} BLAS threads")


def emit(line: dict, fd: int) -> None:
    os.write(fd, (json.dumps(line) + "\n").encode())


def main():
    # stdout carries exactly ONE JSON line: libraries (NCCL prints its version banner) get stderr instead
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--p2p-blocks", type=int, default=0, help="experiment: CTAs per SM of the peer-memory exchange kernel")
    ap.add_argument("--p2p", action="store_true",
                    help="multi-GPU: fused NVLink peer-memory reduce-scatter + update + all-gather kernel instead of ncclAllReduce + "
                         "update (measured slower at 8 GPUs this round, see profiles/README.md)")
    ap.add_argument("--nvls", type=int, default=1,
                    help="multi-GPU, bf16: 1 = fused reduce-scatter + update + all-gather kernel over NVSwitch multicast (multimem.ld_reduce / "
                         "multimem.st) when the box supports it, else the NCCL exchange; 0 = NCCL exchange")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="experiment: override the workload's global batch (e.g. 2048 on 2 GPUs reproduces the 1024-rows-per-rank regime "
                         "of the 8-GPU run); the line's config then names it")
    ap.add_argument("--split-exchange", type=int, default=-1,
                    help="NVLS path, binary PCD: 0 = never split the gradient exchange, 2 = always, -1 = library default (automatic: "
                         "positive product reduced behind the chain advance + negative product as packed uint16 counts when a "
                         "half-step leaves SMs idle)")
    ap.add_argument("--sharded-update", type=int, default=-1,
                    help="multi-GPU exchange: 1 = NCCL reduce-scatter + shard update + bf16 all-gather, 0 = allreduce + full update, -1 = library default")
    ap.add_argument("--state-exchange", type=int, default=-1,
                    help="multi-GPU PCD: 1 = reduce-scatter of the positive product behind the chains + bit-packed chain-state all-gather")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: global batch fixed (BASELINE config); weak: per-GPU batch fixed (global batch x N)")
    ap.add_argument("--stale-chains", action="store_true", help="also time the stale-1 chains mode on one GPU")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--sustained-seconds", type=float, default=3.0,
                    help="extra leg: run steps back to back for this long (own clock record) so that one of the two measured "
                         "bf16 peaks (burst / sustained) is in-regime; 0 disables it")
    ap.add_argument("--no-e2e-i64", action="store_true", help="skip the end-to-end leg on Int64 host rows (Vector{Vector{Int}})")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.global_batch:
        w["B"] = args.global_batch
        w["desc"] += f" [global batch overridden: {args.global_batch}]"
    if args.scaling == "weak":
        w["B"] *= world
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    Bl_cfg = w["B"] // max(world, 1)
    cfg = {"workload": f"{args.workload}: {w['desc']}", "V": w["V"], "H": w["H"], "L": w["L"], "global_batch": w["B"],
           "l2": "inputs cycle over %d resident mini-batches (%.2f GB incl. transposed copies) > 126 MB L2"
                 % (w["nb"], 2 * w["nb"] * Bl_cfg * w["V"] * 2 / 1e9) if 2 * w["nb"] * Bl_cfg * w["V"] * 2 > 126e6 else
                 "inputs cycle over %d resident mini-batches (%.3f GB): smaller than the 126 MB L2, which is the steady state of "
                 "training on this dataset (every epoch re-reads it); no flush between steps" % (w["nb"], 2 * w["nb"] * Bl_cfg * w["V"] * 2 / 1e9),
           "gibbs_steps": w["k"], "algorithm": w["algo"], "rng": "philox4x32-10", "persistent_chains": w.get("chains", w["B"]) if w["algo"] == "PCD" else None,
           "parallelism": f"dp{world} (rows/chains of every mini-batch sharded by rank, one gradient exchange per step)" if world > 1 else "single GPU"}
    # how the gradient exchange is done is a property of THIS arm (decided at run time), reported next to `config`, not inside it
    exchange_mode = ("fused NVLink peer-memory reduce-scatter + update + bf16 all-gather kernel (unicast, CUDA IPC)" if (args.p2p and args.precision == "bf16") else
                     "ncclReduceScatter + shard update + bf16 ncclAllGather + local transpose" if (args.sharded_update != 0 and args.precision == "bf16" and w["H"] % world == 0) else
                     "ncclAllReduce + update") if world > 1 else None

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference(w, 60.0, args.steps, args.warmup)
        rb = cpu_batched_blas(w, 5.0)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "samples/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["threads"], "kind": "port",
                                 "sample": r["sample"], "value_batched_blas": rb["value"], "sample_batched_blas": rb["sample"]},
                "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gibbs_steps_per_s": r["value"] * w["k"] * (w.get("chains", w["B"]) / w["B"] if w["algo"] == "PCD" else 1.0),
                "note": "CPU restatement of QARBoM.jl (C, Float64, per-sample GEMV + outer products; Julia not installable here)"}
        emit(line, out_fd)
        return 0

    import torch
    import torch.distributed as dist
    import qarbom_b200 as q

    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, k, nb = w["B"], w["k"], w["nb"]
    assert B % world == 0 and (B // world) % 4 == 0
    Bl = B // world
    X, Y, W0, U0 = make_data(w, 1000 + int(args.workload[3]))
    Xl, Yl = shard_rows(X, B, nb, rank, world), shard_rows(Y, B, nb, rank, world)
    if w["L"]:
        rbm = (q.GRBMClassifier if w["gaussian"] else q.RBMClassifier)(w["V"], w["H"], w["L"], W=W0, U=U0)
        if w["gaussian"]:
            rbm.a[:] = 0; rbm.b[:] = 0; rbm.c[:] = 0
    else:
        rbm = (q.GRBM if w["gaussian"] else q.RBM)(w["V"], w["H"], W=W0)
    chains_l = w.get("chains", B) // world
    args.p2p_like = False
    want_nvls = (world > 1 and args.precision == "bf16" and bool(args.nvls) and not args.p2p and args.sharded_update != 0 and
                 args.state_exchange <= 0 and w["H"] % world == 0)
    for try_nvls in ((True, False) if want_nvls else (False,)):
        dev = q.DeviceRBM(rbm, max_batch=max(Bl, chains_l), precision=args.precision, device=local_rank, rank=rank, world=world, seed=1234)
        if world == 1:
            break
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(q.DeviceRBM.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        dev.comm_init(bytes(uid.cpu().numpy().tobytes()))
        if args.precision == "bf16" and args.p2p:
            # NVLink peer-memory path: reduce-scatter + update + all-gather of the bf16 shadows in one kernel per rank
            mine = torch.frombuffer(bytearray(dev.p2p_export()), dtype=torch.uint8).cuda()
            allh = torch.empty(320 * world, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(allh, mine)
            dev.p2p_import(allh.cpu().numpy().tobytes())
            if args.p2p_blocks:
                dev.set_option("p2p_blocks_per_sm", args.p2p_blocks)
        elif args.precision == "bf16" and args.sharded_update >= 0 and w["H"] % world == 0:
            dev.set_option("sharded_update", args.sharded_update)
        nvls_on = False
        if try_nvls:
            try:
                why = dev.enable_nvls_exchange()      # collective; "" = switched on
            except q.QbError as e:                    # anything but "unsupported": still only a reason not to use it
                why = str(e)
            # every rank must have switched, or none: agree, and start over on the NCCL exchange otherwise
            ok = torch.tensor([1 if why == "" else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            nvls_on = bool(ok.item())
            if nvls_on:
                exchange_mode = ("ONE fused kernel per rank over NVSwitch multicast: multimem.ld_reduce of the fp32 gradient shard (reduce-scatter in "
                                 "the switch) + update of the owned hidden rows + multimem.st of the new bf16 shadow (all-gather by the switch); "
                                 "transposed shadow rebuilt locally")
            else:
                if why:
                    print(f"[bench] rank {rank}: NVSwitch multicast exchange not available, using the NCCL exchange: {why}", file=sys.stderr)
                dev.close()
                continue
        args.p2p_like = args.p2p or nvls_on
        if nvls_on and args.split_exchange >= 0:
            dev.set_option("split_exchange", args.split_exchange)
        if args.precision == "bf16" and args.state_exchange >= 0 and not args.p2p_like and args.sharded_update != 0 and w["H"] % world == 0:
            dev.set_option("state_exchange", args.state_exchange)
        break
    dev.set_data(Xl, Yl)
    pcd = w["algo"] == "PCD"
    if pcd:
        dev.init_chains(chains_l)
    lr = 0.01
    step = (lambda i: dev.pcd_step(i % nb, Bl, k, lr, lr)) if pcd else (lambda i: dev.cd_step(i % nb, Bl, k, lr, lr))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        dev.sync()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---------------- device-resident leg: pass A (clean, for `value`), pass B (GEMM/update launches
    # bracketed by CUDA events, for the rooflines; the event records perturb tiny kernels and PDL)
    dev.set_option("async_steps", 1)

    def run_steps(first, n):
        """n consecutive mini-batch steps on the resident batches first, first + 1, ... (mod nb).  On one GPU this is ONE call of
        qb_train_steps -- what the epoch calls of the API do: small models then run as one persistent launch
        (sm100_epoch.cuh), everything else as the same per-step launch sequence as qb_pcd_step / qb_cd_step."""
        if world == 1:
            dev.train_steps(w["algo"], first % nb, n, Bl, k, lr, lr)
        else:
            for i in range(n):
                step(first + i)

    run_steps(0, args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    dev.reset_counters()
    if world > 1 and args.precision == "bf16" and args.p2p_like:
        dev.p2p_timing()  # reset the in-kernel phase timers
    dev.timer_start()
    run_steps(args.warmup, args.steps)
    ms = dev.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    launches = dev.counters_full()[0]
    value = B * args.steps / (ms * 1e-3)
    p2p_t = dev.p2p_timing() if (world > 1 and args.precision == "bf16" and args.p2p_like) else None
    dev.reset_counters()
    dev.set_option("time_gemms", 1)
    barrier()
    dev.timer_start()
    for i in range(args.steps):
        step(args.warmup + i)
    ms_instr = max_over_ranks(dev.timer_stop())
    barrier()
    _, gemms, gflops, timed, gemm_ms, timed_flops = dev.counters_full()
    upd_n, upd_ms = dev.update_timing()
    dev.set_option("time_gemms", 0)

    # ---------------- sustained leg: >= `--sustained-seconds` of back-to-back steps with every GEMM bracketed by events and its
    # own clock record: the regime the sustained bf16 peak of MEASURED_PEAKS.json was measured in (the K-step region above is
    # a burst of a few hundred ms at most).  Event bracketing is only meaningful for long kernels; tiny configs report the
    # whole-step figure alone.
    sustained = None
    if args.sustained_seconds > 0 and args.precision == "bf16":
        n_sus = int(min(20000, max(args.steps, np.ceil(args.sustained_seconds * 1e3 / max(ms / args.steps, 1e-3)))))
        big = (gemm_ms / max(timed, 1)) > 0.05   # average GEMM launch longer than 50 us
        dev.reset_counters()
        dev.set_option("time_gemms", 1 if big else 0)
        barrier()
        sampler2 = ClockSampler(local_rank)
        if rank == 0:
            sampler2.start()
        dev.timer_start()
        run_steps(args.warmup, n_sus)
        ms_sus = max_over_ranks(dev.timer_stop())
        barrier()
        clocks_sus = sampler2.stop() if rank == 0 else None
        _, _, _, timed_s, gemm_ms_s, timed_flops_s = dev.counters_full()
        dev.set_option("time_gemms", 0)
        sustained = {"steps": n_sus, "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus, "value": B * n_sus / (ms_sus * 1e-3),
                     "unit": "samples/s", "clocks": clocks_sus,
                     "gemm_tflops": (timed_flops_s / 1e12) / (gemm_ms_s * 1e-3) if (big and gemm_ms_s > 0) else None,
                     "gemm_launches_timed": timed_s if big else 0}

    # ---------------- optional leg: declared stale-1 chains mode (allreduce + update overlapped with the next
    # step's chain advance; statistical parity only)
    overlap = None
    if pcd and args.precision == "bf16" and w["L"] == 0 and ((world > 1 and (not args.p2p_like)) or args.stale_chains):
        dev.set_option("stale_chains", 1)
        for i in range(args.warmup):
            step(i)
        barrier()
        dev.timer_start()
        for i in range(args.steps):
            step(args.warmup + i)
        ms_st = max_over_ranks(dev.timer_stop())
        barrier()
        dev.set_option("stale_chains", 0)
        overlap = {"mode": "stale-1 chains: allreduce+update of step t overlap the chain advance of step t+1",
                   "value": B * args.steps / (ms_st * 1e-3), "unit": "samples/s", "ms_per_step": ms_st / args.steps,
                   "semantics": "declared extension, not used for `value`"}

    # ---------------- end-to-end leg: host mini-batches in pinned memory, recon error read back
    e2e = None
    if not args.no_e2e:
        nbuf = min(nb, 4)
        xs = [torch.from_numpy(Xl[j * Bl:(j + 1) * Bl]).pin_memory() for j in range(nbuf)]
        ys = [torch.from_numpy(Yl[j * Bl:(j + 1) * Bl]).pin_memory() for j in range(nbuf)] if Yl is not None else None
        dev.set_option("async_steps", 0)
        def host_step(i):  # the next mini-batch is uploaded (pinned -> HBM) while this one is being processed
            j, jn = i % nbuf, (i + 1) % nbuf
            return dev.step_host(w["algo"], xs[j], ys[j] if ys else None, k, lr, lr, next_x=xs[jn],
                                 next_y=ys[jn] if ys else None)

        for i in range(args.warmup):
            host_step(i)
        barrier()
        dev.timer_start()
        for i in range(args.steps):
            host_step(args.warmup + i)
        ms2 = max_over_ranks(dev.timer_stop())
        barrier()
        h2d = int(xs[0].numel() * xs[0].element_size() + (ys[0].numel() * ys[0].element_size() if ys else 0))
        e2e = {"value": B * args.steps / (ms2 * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": 8 * world, "ms_per_step": ms2 / args.steps,
               "api": f"qb_{'pcd' if pcd else 'cd'}_step_host + qb_set_next_host_batch (pinned-host upload of every mini-batch, "
                      "overlapped with the previous step; step; reconstruction error readback)"}

    # ---------------- the same end-to-end leg on the reference API's own host element type: Vector{Vector{Int}} = Int64 rows
    # (Float64 for Gaussian data), 8 bytes per element over PCIe instead of 1.  The Julia shim packs x_train to UInt8 ONCE per
    # train! call (QARBoMB200.jl, `pack_rows`), so `e2e` above is what an epoch loop sees; this is the cost without packing.
    e2e_i64 = None
    if e2e is not None and not args.no_e2e_i64:
        wide = np.float64 if w["gaussian"] else np.int64
        nbuf = min(nb, 2)
        xs = [torch.from_numpy(Xl[j * Bl:(j + 1) * Bl].astype(wide)).pin_memory() for j in range(nbuf)]
        ys = [torch.from_numpy(Yl[j * Bl:(j + 1) * Bl].astype(np.int64)).pin_memory() for j in range(nbuf)] if Yl is not None else None

        def host_step64(i):
            j, jn = i % nbuf, (i + 1) % nbuf
            return dev.step_host(w["algo"], xs[j], ys[j] if ys else None, k, lr, lr, next_x=xs[jn], next_y=ys[jn] if ys else None)

        n64 = max(3, min(args.steps, 10))
        for i in range(3):
            host_step64(i)
        barrier()
        dev.timer_start()
        for i in range(n64):
            host_step64(3 + i)
        ms3 = max_over_ranks(dev.timer_stop())
        barrier()
        h2d = int(xs[0].numel() * xs[0].element_size() + (ys[0].numel() * ys[0].element_size() if ys else 0))
        e2e_i64 = {"value": B * n64 / (ms3 * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 8 * world,
                   "ms_per_step": ms3 / n64, "steps": n64, "host_dtype": "float64" if w["gaussian"] else "int64",
                   "pcie_gbs": h2d / (ms3 / n64 * 1e-3) / 1e9}
        del xs, ys

    if rank != 0:
        dev.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    pk = peaks()
    # Two measured bf16 peaks exist (MEASURED_PEAKS.json): burst (a kernel timed alone at full clocks) and sustained (back to
    # back for seconds, power-capped clocks).  The K-step timed region is short, so `frac` divides by the BURST peak unless the
    # clock samples of that region show the power cap; `frac_burst` / `frac_sustained` are always both printed, and the
    # sustained leg below has its own achieved figure measured in the sustained regime.
    reasons = (clocks or {}).get("reasons") or []
    sm_med, sm_max = (clocks or {}).get("sm_mhz") or 0, (clocks or {}).get("sm_max_mhz") or 1
    capped = ("sw_power_cap" in reasons and sm_med < 0.9 * sm_max) or (sm_med and sm_med < 0.8 * sm_max)
    tf_burst, tf_sust = pk.get("tf_burst", pk["tf"]), pk["tf"]
    pk = dict(pk, tf=tf_sust if capped else tf_burst,
              src=pk["src"].replace("sustained bf16", "sustained bf16: timed region ran power-capped" if capped else
                                    "burst bf16: timed region ran at >= 90 % of max SM clock"))
    F = flops_per_step(w) / world  # per GPU per step
    achieved = (timed_flops / 1e12) / (gemm_ms * 1e-3) if gemm_ms > 0 else None
    n_sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    cfg5_bf16 = args.workload == "cfg5" and args.precision == "bf16" and world == 1
    roof = {"bound": "tensor", "achieved": achieved, "peak": pk["tf"], "unit": "TFLOP/s",
            "frac": (achieved / pk["tf"]) if achieved else None,
            "frac_burst": (achieved / tf_burst) if achieved else None, "frac_sustained": (achieved / tf_sust) if achieved else None,
            "peak_burst": tf_burst, "peak_sustained": tf_sust,
            # against the hardware instead of a library: 148 SMs x 8192 dense bf16 flop per clock x the median SM clock of the region
            "frac_of_pipe_at_clock": (achieved / (n_sms * 8192 * sm_med * 1e6 / 1e12)) if (achieved and sm_med) else None,
            "traffic": ncu_traffic(NCU_GEMM_SUMMARY, "gibbs_", mix=gemm_mix(w)) if cfg5_bf16 else None,
            "traffic_note": "dram read+write bytes per launch from one `ncu --set full` capture of this command (profiles/" + NCU_GEMM_SUMMARY +
                            "), per-kernel means weighted by the launch mix of a step (1 chain launch = the 2k + 1 forward contractions : 1 gradient GEMM)",
            "kernel": "qb::sm100::gibbs_chain_kernel (persistent: all fused Gibbs half-steps of the step) + gibbs_gemm_kernel_2cta<., K_GRAD> (gradient); "
                      "gibbs_gemm_kernel* per half-step where the chain kernel does not pay" if args.precision == "bf16"
            else "qb::simt_gemm_kernel (fp32 validation path)",
            "gemm_launches_timed": timed, "avg_launch_ms": gemm_ms / max(timed, 1),
            "gemm_share_of_step": gemm_ms / ms_instr if ms_instr > 0 else None,
            "instrumented_ms_per_step": ms_instr / args.steps,
            "algorithmic_flops_per_step_per_gpu": F, "issued_flops_per_step_per_gpu": gflops / args.steps,
            "peak_source": pk["src"],
            "peak_note": "the denominator is a MEASURED cuBLAS bf16 8192^3 throughput (MEASURED_PEAKS.json), not the nominal 2250 TFLOP/s; "
                         "a fraction slightly above 1 means these GEMMs ran faster than cuBLAS did in the same power regime",
            "whole_step_tflops": F * args.steps / (ms * 1e-3) / 1e12}
    if sustained:
        ws = F * sustained["steps"] / sustained["seconds"] / 1e12
        sustained.update({"whole_step_tflops": ws, "whole_step_frac_sustained": ws / tf_sust,
                          "gemm_frac_sustained": (sustained["gemm_tflops"] / tf_sust) if sustained["gemm_tflops"] else None,
                          "gemm_frac_burst": (sustained["gemm_tflops"] / tf_burst) if sustained["gemm_tflops"] else None})
        roof["sustained"] = sustained
    # update kernel (HBM-bound): 12 B/param fp32 read-modify-write + 4 B/param for the two bf16 shadows
    # (read w, read g, write w); with the NCCL sharded update a rank touches only its H / world hidden rows and writes only the
    # row-major shadow (the transposed one is rebuilt after the all-gather by transpose_bf16_wide_kernel)
    P_w = (w["V"] + w["L"]) * w["H"]
    sharded = world > 1 and args.precision == "bf16" and not args.p2p_like and args.sharded_update != 0 and w["H"] % world == 0
    upd_bytes = (16.0 if args.precision == "bf16" else 12.0) * P_w
    if sharded:
        upd_bytes = 14.0 * P_w / world
    upd_gbs = (upd_bytes * upd_n / 1e9) / (upd_ms * 1e-3) if upd_ms > 0 else None
    roof_upd = {"bound": "hbm", "achieved": upd_gbs, "peak": pk["hbm"], "unit": "GB/s",
                "frac": (upd_gbs / pk["hbm"]) if upd_gbs else None,
                "traffic": ncu_traffic(NCU_UPDATE_SUMMARY, "update_w_kernel") if cfg5_bf16 else None,
                "bytes_note": "14 B/param x (H/world) rows: sharded update" if sharded else "12 B/param fp32 read-modify-write + 4 B/param bf16 shadows",
                "kernel": "qb::update_w_kernel",
                "bytes_per_launch": upd_bytes, "avg_launch_ms": upd_ms / max(upd_n, 1), "launches_timed": upd_n}
    if world > 1 and args.p2p_like and not args.p2p:
        # the "update" launch of this mode is the fused exchange kernel (reduce-scatter + shard update + all-gather over NVLink)
        # plus the local transpose: bound by the link, not by HBM.  Bytes a GPU sends per step: its whole fp32 gradient leaves it
        # once (every shard is summed elsewhere -- through the switch also its own) + its shard of the new bf16 shadow.
        nvl_out = 4.0 * P_w + 2.0 * P_w / world
        gbs = (nvl_out * upd_n / 1e9) / (upd_ms * 1e-3) if upd_ms > 0 else None
        roof_upd = {"bound": "nvlink", "achieved": gbs, "peak": 900.0, "unit": "GB/s", "frac": (gbs / 900.0) if gbs else None,
                    "traffic": None, "peak_source": "nominal NVLink 5 bandwidth per direction per GPU (no measured figure in MEASURED_PEAKS.json)",
                    "bytes_note": "outbound NVLink bytes per GPU per step: 4 B/param (fp32 gradient, read by the switch) + 2 B/param / world (bf16 shard pushed); "
                                  "inbound is smaller (4 B/param / world + 2 B/param)",
                    "kernel": "qb::p2p_reduce_update_kernel<., NVLS> + transpose_bf16_wide_kernel",
                    "bytes_per_launch": nvl_out, "avg_launch_ms": upd_ms / max(upd_n, 1), "launches_timed": upd_n}
    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": cfg,
            "gibbs_steps_per_s": value * k * (w.get("chains", B) / B if pcd else 1.0), "gpu_launches": launches, "roofline": roof, "roofline_update": roof_upd, "clocks": clocks, "e2e": e2e}
    # latency floor of the step: its dependent stages (the 2k (+1) forward contractions of the chain, then gradient + update) --
    # on the small configurations a stage is a fraction of a microsecond of MMA and the step time is stages x stage latency
    stages = 2 * k + 2
    line["latency"] = {"dependent_stages_per_step": stages, "us_per_stage": ms / args.steps * 1e3 / stages,
                       "mma_floor_us_per_step": F / (tf_burst * 1e12) * 1e6,
                       "note": "stages = 2k + 1 forward contractions on the critical path (PCD: the positive phase runs beside the chain) + gradient/update"}
    if exchange_mode:
        line["exchange_mode"] = exchange_mode
    if e2e_i64:
        line["e2e_i64"] = e2e_i64
    if overlap:
        line["overlap"] = overlap
    if p2p_t:
        line["exchange"] = {"kernel": "qb::p2p_reduce_update_kernel", "wait_for_peers_ms_per_step": p2p_t[0] / args.steps,
                            "reduce_update_push_ms_per_step": p2p_t[1] / args.steps,
                            "nvlink_bytes_per_step_per_gpu": 2 * 4.0 * (w["V"] + w["L"]) * w["H"] * (world - 1) / world}
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline is a single-GPU-run item (rank 0 at N = 1 only)
        r = cpu_reference(w, args.cpu_seconds, 3, 1)
        r1 = cpu_reference(w, args.cpu_seconds / 2, 2, 1, threads=1)   # the Julia code itself is single-threaded
        rb = cpu_batched_blas(w, args.cpu_seconds / 3)                 # best case for a CPU: batched BLAS-3 form (not the reference's shape)
        line["cpu_baseline"] = {"value": r["value"], "unit": "samples/s", "cores": r["threads"], "kind": "port",
                                "sample": r["sample"], "value_1_thread": r1["value"], "sample_1_thread": r1["sample"],
                                "value_batched_blas": rb["value"], "sample_batched_blas": rb["sample"]}
    emit(line, out_fd)
    dev.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
