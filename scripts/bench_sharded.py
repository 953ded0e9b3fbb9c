#!/usr/bin/env python
"""configs[4] of BASELINE.json: one large state, C1-style random circuit,
depth 20, expectation of sum Z_i.  Under torchrun the state is sharded over
the ranks (NCCL all-to-all qubit swaps); with one process it is the plain op.

  python scripts/bench_sharded.py --qubits 30
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
      scripts/bench_sharded.py --qubits 36 [--check]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_b200 import circuits as cq  # noqa: E402
from quantum_b200 import ops, sharded  # noqa: E402


def grid_shape(n):
    r = int(np.floor(np.sqrt(n)))
    while n % r:
        r -= 1
    return r, n // r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--warmups", type=int, default=1,
                    help="untimed calls first (a single-row plan is specialised "
                         "from its second call on)")
    ap.add_argument("--check", action="store_true",
                    help="compare with the unsharded op on rank 0 (needs the "
                         "whole state on one GPU)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="peer: qubit swaps inside the library through peer "
                         "memory (CUDA IPC over NVLink); nccl: host-driven "
                         "all_to_all_single, the library baseline")
    ap.add_argument("--xterms", action="store_true",
                    help="add sum X_i (forces a qubit swap in the expectation)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rows, cols = grid_shape(a.qubits)
    m, qs = cq.supremacy_style_circuit(rows, cols, a.depth, a.qubits, use_line=True)
    prog = cq.serialize(m)
    terms = [(1.0, [(q, "Z")]) for q in qs]
    if a.xterms:
        terms += [(1.0, [(q, "X")]) for q in qs]
    sums = [cq.pauli_sum(terms)]
    vals = np.zeros(0, np.float32)
    import torch
    torch.cuda.set_device(local)
    stats = {}
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if a.exchange == "nccl":
            run = lambda: sharded.sharded_expectation(prog, [], vals, sums,
                                                      device=local, stats=stats)
        else:
            job = sharded.peer_sharded_job(prog, [], vals, sums, device=local)

            def run():
                job.enqueue()
                out = job.result()
                stats.update(job.stats(), stages=list(job.kinds))
                return out
    else:
        run = lambda: ops.tfq_simulate_expectation(
            [prog], [], np.zeros((1, 0), np.float32), [sums], device=local)[0]
    for _ in range(max(a.warmups, 1)):    # warm-up (plans, kernels, NCCL channels)
        out = run()
    times = []
    for _ in range(a.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = run()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt.item())
    ref = None
    if a.check and rank == 0:
        ref = float(ops.tfq_simulate_expectation(
            [prog], [], np.zeros((1, 0), np.float32), [sums], device=local)[0, 0])
    if rank == 0:
        line = {"config": "C5 single %d-qubit state, depth %d, sum Z%s"
                          % (a.qubits, a.depth, " + sum X" if a.xterms else ""),
                "n_gpus": world, "seconds_per_circuit": t,
                "expectation": float(np.ravel(out)[0]), "unsharded_check": ref,
                "state_gib": 8 * 2 ** a.qubits / 2 ** 30}
        if stats and "pull_ms" in stats:
            ex = stats["exchanges"]
            line.update(exchange="peer memory (in library)", exchanges=ex,
                        fused_into_next_pass=stats.get("fused_exchanges", 0),
                        exchange_seconds=stats["pull_ms"] * 1e-3,
                        wait_seconds=stats["wait_ms"] * 1e-3, stages=stats["stages"],
                        gate_passes=stats["gate_passes"])
            if ex and stats["pull_ms"] > 0:
                line["nvlink_recv_GBps_per_gpu"] = (
                    stats["bytes_received_per_exchange"] * ex / (stats["pull_ms"] * 1e-3) / 1e9)
        elif stats:
            ex = stats.get("exchanges", 0)
            line.update(exchange="nccl all_to_all_single (host driven)", exchanges=ex,
                        exchange_seconds=stats["exchange_seconds"],
                        stages=stats["stages"])
            if ex and stats["exchange_seconds"] > 0:
                # each rank sends (1 - 1/world) of its shard per exchange
                sent = stats["shard_bytes"] * (1 - 1 / world) * ex
                line["nvlink_send_GBps_per_gpu"] = sent / stats["exchange_seconds"] / 1e9
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()          # nobody unmaps a shard a peer may still read
        if a.exchange == "peer":
            job.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
