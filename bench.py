#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json's north_star.

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (configs[1]): 20-qubit hardware-efficient ansatz (4 layers, 160
symbols), batch 4096 rows PER GPU (weak scaling), 4 PauliSum observables.
One step = one tfq_simulate_expectation pass over the batch; the adjoint
gradient (tfq_adj_grad) over the same batch is timed as a second leg and
reported under "adjoint".

`value`   = circuit evaluations / s, device-timed, inputs resident in HBM.
`e2e`     = the same through the public API (quantum_b200.ops, i.e. the C
            ABI with host buffers: parse + plan + H2D + kernels + D2H).
`roofline`= the forward gate-pass kernel: algorithmic 16*2^n B per state per
            pass over its CUDA-event launch time, against MEASURED_PEAKS.json.
`--impl reference`: the restated qsim-style CPU path (oracle/, C executor,
one thread per circuit like ComputeSmall) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_QUBITS = 20
LAYERS = 4
BATCH = 4096
METRIC = "circuit-evals/sec (20q batched expectation)"
UNIT = "circuits/s"
FALLBACK_HBM_GBS = 6650.0


def workload(batch, n=N_QUBITS, layers=LAYERS, seed=20):
    from quantum_b200 import circuits as cq
    moments, names, qs = cq.hea_circuit(n, layers)
    prog = cq.serialize(moments)
    obs = cq.hea_observables(qs)
    rng = np.random.default_rng(seed)
    vals = rng.uniform(0, 2, (batch, len(names))).astype(np.float32)
    down = np.ones((batch, len(obs)), np.float32)
    return prog, names, obs, vals, down


def config_dict(n_gpus, batch):
    return {"workload": "configs[1]: %d-qubit hardware-efficient ansatz, "
                        "%d layers, %d symbols, batch %d per GPU, 4 PauliSum "
                        "observables, tfq_simulate_expectation (+ tfq_adj_grad "
                        "leg)" % (N_QUBITS, LAYERS, 2 * N_QUBITS * LAYERS, batch),
            "n_qubits": N_QUBITS, "batch_per_gpu": batch,
            "global_batch": batch * n_gpus, "n_ops": 4,
            "sharding": "batch rows over ranks, no collective",
            "l2": "inputs larger than L2 (%.0f GiB of states per step)"
                  % (batch * 8 * 2 ** N_QUBITS / 2 ** 30)}


# --------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(
                self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------
def cpu_port_time(n_circuits, threads, adjoint=False):
    """Restated qsim-style CPU path (oracle, C executor) on `n_circuits` rows
    of the workload, `threads` circuits at a time. Returns seconds."""
    from oracle import tfq_oracle as orc
    prog, names, obs, vals, down = workload(n_circuits)
    t0 = time.perf_counter()
    if adjoint:
        orc.adjoint_gradient([prog] * n_circuits, names, vals,
                             [obs] * n_circuits, down, threads=threads)
    else:
        orc.simulate_expectation([prog] * n_circuits, names, vals,
                                 [obs] * n_circuits, threads=threads)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 4 * cores  # circuits per step, one per host thread at a time (ComputeSmall)
    from oracle import tfq_oracle as orc
    orc.build_c()
    for _ in range(max(args.warmup, 0) and 1):
        cpu_port_time(sample, cores)
    times = [cpu_port_time(sample, cores) for _ in range(max(args.steps, 1))]
    t = float(np.sum(times))
    value = sample * len(times) / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 1),
        "ms_per_step": 1e3 * t / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "complex64",
        "data": "synthetic", "config": config_dict(args.gpus, BATCH),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores,
                         "kind": "port",
                         "sample": "%d circuits of the workload per step (one "
                                   "per host thread at a time), restated qsim-style CPU "
                                   "path (oracle/qsim_vm.c); qsim itself is "
                                   "not installable here" % sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-adjoint", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    import torch
    import torch.distributed as dist
    from quantum_b200 import ops

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch
    prog, names, obs, vals, down = workload(B, seed=20 + rank)
    programs = [prog] * B
    sums = [obs] * B
    ctx = ops.get_context(local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local)

    def timed(fn, steps, profile=False):
        """K steps bracketed by barrier + sync; CUDA events on the library's
        stream; returns (seconds max over ranks, profile dict)."""
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        barrier()
        ctx.sync()
        torch.cuda.synchronize()
        ctx.profile_reset()
        ctx.profile_enable(profile)
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        barrier()
        sec = ev0.elapsed_time(ev1) * 1e-3
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        return max_over_ranks(sec), prof

    # ---- leg 1: expectation, device resident
    job = ops.DeviceJob("expectation", programs, names, vals, sums, device=local)
    for _ in range(W):
        job.run()
    ctx.sync()
    sampler = ClockSampler(local)
    sampler.start()
    sec, prof = timed(job.run, K, profile=True)
    clocks = sampler.finish()
    result_dev = job.fetch()
    job.close()
    value = world * B * K / sec

    peak, peak_src = measured_peak()
    gp_launches = max(prof["gate_pass_launches"], 1)
    gp_ms = prof["gate_pass_ms"]
    achieved = prof["gate_pass_bytes"] / max(gp_ms * 1e-3, 1e-12) / 1e9
    traffic = None
    try:   # dram__bytes_read+write per launch from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("n_qubits") == N_QUBITS:
            n_pass = prof["gate_pass_launches"] / max(K, 1)   # passes per step
            per_state = (tj["forward_first_pass_dram_bytes_per_state"] +
                         (n_pass - 1) * tj["forward_pass_rw_dram_bytes_per_state"]) / n_pass
            traffic = per_state * B
    except Exception:
        traffic = None
    roofline = {
        "bound": "hbm",
        "kernel": ("tfqb_jit_pass (forward gate pass, specialised at run time from "
                   "pass_device.cuh by csrc/jit.cc; FP32-pipe limited, see DESIGN.md 6)"
                   if prof.get("jit_pass_launches", 0) > 0 else
                   "pass_kernel<4,2,false> (interpreted forward gate pass; "
                   "FP32-pipe limited, see DESIGN.md 6)"),
        "specialised_launches": int(prof.get("jit_pass_launches", 0)),
        "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic,
        "traffic_source": "ncu dram__bytes_read.sum+dram__bytes_write.sum per launch, "
                          "profiles/r01_traffic.json (from r01j_forward_expect_full_raw.csv), scaled to this batch",
        "peak_source": peak_src,
        "launches": int(prof["gate_pass_launches"]),
        "avg_launch_ms": gp_ms / gp_launches,
        "algorithmic_bytes_per_launch": prof["gate_pass_bytes"] / gp_launches,
        "share_of_step": gp_ms * 1e-3 / sec,
        "expectation_kernel": {
            "achieved": prof["expectation_bytes"] /
            max(prof["expectation_ms"] * 1e-3, 1e-12) / 1e9,
            "unit": "GB/s", "launches": int(prof["expectation_launches"]),
            "share_of_step": prof["expectation_ms"] * 1e-3 / sec},
    }
    launches = int(prof["kernel_launches"])

    # ---- leg 2: end to end through the public API (host buffers)
    e2e_steps = max(1, min(K, 10))
    ops.tfq_simulate_expectation(programs, names, vals, sums, device=local)
    barrier()
    ctx.profile_reset()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        result_e2e = ops.tfq_simulate_expectation(programs, names, vals, sums,
                                                  device=local)
    e2e_sec = max_over_ranks(time.perf_counter() - t0)
    prof_e2e = ctx.profile_read()
    assert np.allclose(result_e2e, result_dev, atol=1e-6)
    e2e = {"value": world * B * e2e_steps / e2e_sec, "unit": UNIT,
           "h2d_bytes_per_step": int(prof_e2e["h2d_bytes"] // e2e_steps),
           "d2h_bytes_per_step": int(prof_e2e["d2h_bytes"] // e2e_steps),
           "steps": e2e_steps,
           "note": "ops.tfq_simulate_expectation(host strings + float32 "
                   "arrays) -> numpy; includes proto parse, planning, H2D, D2H"}

    # ---- leg 3: adjoint gradient
    adjoint = None
    if not args.no_adjoint:
        ajob = ops.DeviceJob("adjoint", programs, names, vals, sums, down,
                             device=local)
        for _ in range(W):
            ajob.run()
        ctx.sync()
        asec, aprof = timed(ajob.run, K, profile=True)
        adj_result = ajob.fetch()
        ajob.close()
        a_ms = aprof["adjoint_pass_ms"]
        a_ach = aprof["adjoint_pass_bytes"] / max(a_ms * 1e-3, 1e-12) / 1e9
        t0 = time.perf_counter()
        ops.tfq_adj_grad(programs, names, vals, sums, down, device=local)
        a_e2e = max_over_ranks(time.perf_counter() - t0)
        adjoint = {
            "metric": "adjoint-grad circuits/sec", "value": world * B * K / asec,
            "unit": UNIT, "ms_per_step": 1e3 * asec / K,
            "e2e": {"value": world * B / a_e2e, "unit": UNIT},
            "roofline": {"bound": "hbm",
                         "kernel": ("tfqb_jit_pass (fused reverse pass, specialised "
                                    "at run time)"
                                    if aprof.get("jit_pass_launches", 0) > 0 else
                                    "pass_kernel<3,1,true> (fused reverse pass)"),
                         "achieved": a_ach, "peak": peak, "unit": "GB/s",
                         "frac": a_ach / peak,
                         "launches": int(aprof["adjoint_pass_launches"]),
                         "share_of_step": a_ms * 1e-3 / asec},
            "gpu_launches": int(aprof["kernel_launches"])}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same rows
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_c = 16 * cores
        t = cpu_port_time(n_c, cores)
        # the oracle as the checker of the rows that were just timed: the first
        # rows of the device-resident result against the CPU restatement
        from oracle import tfq_oracle as orc
        n_chk = min(8, B)
        ref = orc.simulate_expectation(programs[:n_chk], names, vals[:n_chk],
                                       sums[:n_chk], threads=min(cores, n_chk))
        err = float(np.abs(result_dev[:n_chk] - ref).max())
        parity = {"rows_checked": n_chk, "max_abs_err": err,
                  "tolerance": "1e-5 abs + 1e-4 rel (north_star)",
                  "ok": bool(np.allclose(result_dev[:n_chk], ref, atol=1e-5, rtol=1e-4))}
        if adjoint is not None:
            # the rows of the adjoint leg that was just timed (specialised
            # reverse passes) against the oracle's 12-sweep adjoint step
            n_a = min(4, B)
            gref = orc.adjoint_gradient(programs[:n_a], names, vals[:n_a], sums[:n_a],
                                        down[:n_a], threads=min(cores, n_a))
            parity["adjoint"] = {
                "rows_checked": n_a,
                "max_abs_err": float(np.abs(adj_result[:n_a] - gref).max()),
                "grad_scale": float(np.abs(gref).max()),
                "ok": bool(np.allclose(adj_result[:n_a], gref, atol=1e-5, rtol=1e-4))}
        cpu = {"value": n_c / t, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d circuits of the workload, one per host thread at a time, "
                         "%.1f s (restated qsim-style CPU path, oracle/qsim_vm.c)"
                         % (n_c, t)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * sec / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "complex64", "data": "synthetic",
            "config": config_dict(world, B), "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu, "parity_vs_oracle": parity, "adjoint": adjoint,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
