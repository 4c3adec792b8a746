"""torchrun worker for the data-parallel parity check (tests/test_multi_gpu.py):
G ranks, each with B/G rows and chains of every mini-batch and one NCCL gradient allreduce
per step, must reproduce the single-GPU run (Philox is keyed by the GLOBAL row)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qarbom_b200 as q  # noqa: E402
from bench import shard_rows  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    exchange = sys.argv[2] if len(sys.argv) > 2 else "nccl"   # "p2p": NVLink peer-memory reduce+update kernel; "sharded": NCCL RS + AG
    rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr_)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
    V, H, B, k, nb, steps = 320, 192, 64 * world, 2, 3, 3
    rng = np.random.default_rng(0)
    X = rng.integers(0, 2, (nb * B, V)).astype(np.uint8)
    W0 = 0.1 * rng.standard_normal((V, H))
    Bl = B // world

    def run(rk, wd, Xs, Brows, comm):
        rbm = q.RBM(V, H, W=W0)
        dev = q.DeviceRBM(rbm, max_batch=Brows, precision=prec, device=lr_, rank=rk, world=wd, seed=99)
        if comm:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(q.DeviceRBM.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            dev.comm_init(uid.cpu().numpy().tobytes())
            if prec == "bf16":   # "sharded" (the bf16 default): NCCL reduce-scatter + shard update + bf16 all-gather + local transpose
                dev.set_option("sharded_update", 0 if exchange == "nccl" else 1)   # ("chain" runs on the default sharded exchange)
            if exchange == "state":     # positive product reduce-scattered behind the chains + bit-packed chain-state all-gather
                dev.set_option("state_exchange", 1)
            if exchange in ("nvls", "split"):   # the fused kernel over NVSwitch multicast addresses (NCCL symmetric windows); "split":
                                                # positive product reduced behind the chain advance, negative product as packed uint16 counts
                why = dev.enable_nvls_exchange()
                if why:
                    print(f"MGPU-SKIP nvls: {why}")
                    dev.close()
                    dist.barrier(); dist.destroy_process_group(); sys.exit(77)
                dev.set_option("split_exchange", 2 if exchange == "split" else 0)
            if exchange == "p2p":
                mine = torch.frombuffer(bytearray(dev.p2p_export()), dtype=torch.uint8).cuda()
                allh = torch.empty(320 * world, dtype=torch.uint8, device="cuda")
                dist.all_gather_into_tensor(allh, mine)
                dev.p2p_import(allh.cpu().numpy().tobytes())
        if exchange == "chain" and comm:   # the persistent multi-half-step kernel on the data-parallel side, one launch per half-step on the reference side
            dev.set_option("fused_chain", 2)
        dev.set_data(Xs)
        dev.init_chains(Brows)
        if exchange == "stale":      # stale-1 chains on both sides: the overlapped exchange (CTA-limited communicator) must not change results
            dev.set_option("stale_chains", 1)
        for t in range(steps):
            dev.pcd_step(t % nb, Brows, k, 0.05)
        if comm:
            dist.barrier()   # pull_params gathers the sharded master in p2p mode: every rank must be between steps
        dev.pull_params()
        if comm:
            dist.barrier()
        v, h, _ = dev.get_chains()
        dev.close()
        return rbm.W.copy(), rbm.a.copy(), rbm.b.copy(), v, h

    Wm, am, bm, vm, hm = run(rank, world, shard_rows(X, B, nb, rank, world), Bl, True)
    gathered = [None] * world
    dist.all_gather_object(gathered, (Wm, vm, hm))
    ok = True
    if rank == 0:
        Ws, as_, bs, vs, hs = run(0, 1, X, B, False)
        for r, (Wr, _, _) in enumerate(gathered):
            assert np.array_equal(Wr, gathered[0][0]), f"replica {r} diverged from replica 0"
        rel = np.linalg.norm(Wm - Ws) / np.linalg.norm(Ws)
        drel = np.linalg.norm((Wm - W0) - (Ws - W0)) / np.linalg.norm(Ws - W0)
        vcat = np.concatenate([g[1] for g in gathered]); hcat = np.concatenate([g[2] for g in gathered])
        mis = float(np.mean(vcat != vs)), float(np.mean(hcat != hs))
        print(f"MGPU world={world} prec={prec} exchange={exchange} relW={rel:.3e} rel_dW={drel:.3e} state_mismatch={mis}")
        ok = rel < 1e-5 and drel < (2e-2 if prec == "bf16" else 1e-4) and max(mis) < 1e-3
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
