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
`--impl reference`: the restated qsim-style CPU path (oracle/, C executor with
AVX2 gate loops, one thread per circuit like ComputeSmall) on the host cores;
the timed quantity is the C simulation alone, the oracle's Python preparation
is reported next to it.

Extra legs (keys of the same JSON line):
  `c4_strong`         configs[3]: 22-qubit TFI ansatz adjoint gradients, FIXED
                      global batch split over the ranks (strong scaling);
  `sharded_state`     configs[4]: one state of 34 (N=1) / 33+log2(N) qubits
                      sharded over the ranks, qubit swaps through peer memory
                      inside the library, with a 24-qubit check against the
                      unsharded op;
  `one_call_all_gpus` (N>1) ONE tfq_simulate_expectation call from rank 0 over
                      a multi-device context spanning all N GPUs.
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
TRAFFIC_FILE = next((f for f in ("r03_traffic.json", "r02_traffic.json")
                     if os.path.exists(os.path.join(ROOT, "profiles", f))),
                    "r01_traffic.json")
FFMA2_PER_S = 70.4e12 / 4.0      # measured: scripts/micro/pipe_rates.cu on B200


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
    of the workload, `threads` circuits at a time.  Returns (simulation
    seconds = the C executor alone, preparation seconds = the oracle's Python:
    proto parse, gate matrices, step lists -- work the reference does in C++
    and that is NOT charged to the CPU arm)."""
    from oracle import tfq_oracle as orc
    prog, names, obs, vals, down = workload(n_circuits)
    t0 = time.perf_counter()
    if adjoint:
        orc.adjoint_gradient([prog] * n_circuits, names, vals,
                             [obs] * n_circuits, down, threads=threads)
    else:
        orc.simulate_expectation([prog] * n_circuits, names, vals,
                                 [obs] * n_circuits, threads=threads)
    wall = time.perf_counter() - t0
    run = float(orc.LAST_TIMING["run_s"])
    return run, wall - run


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 8 * cores  # circuits per step, one per host thread at a time (ComputeSmall)
    from oracle import tfq_oracle as orc
    orc.build_c()
    for _ in range(max(args.warmup, 0) and 1):
        cpu_port_time(sample, cores)
    runs = [cpu_port_time(sample, cores) for _ in range(max(args.steps, 1))]
    times = [r[0] for r in runs]
    prep = float(np.sum([r[1] for r in runs]))
    t = float(np.sum(times))
    value = sample * len(times) / t
    # the same sample on half the threads: does the arm scale with cores?
    half = max(1, cores // 2)
    t_half = cpu_port_time(sample, half)[0]
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
                                   "path (oracle/qsim_vm.c, AVX2 gate loops); timed: the C "
                                   "simulation alone; qsim itself is "
                                   "not installable here" % sample,
                         "python_prep_s_not_charged": prep,
                         "value_on_half_the_cores": sample / t_half,
                         "half_cores": half},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)



# --------------------------------------------------------------------------
# extra legs
# --------------------------------------------------------------------------
def leg_c4_strong(args, ops, ctx, rank, world, local, timed, max_over_ranks, peak):
    """configs[3]: 22-spin TFI ansatz (spin_system.py:254-261), 22 symbols over
    484 parameterised gates, tfq_adj_grad; a FIXED global batch split over the
    ranks in contiguous row blocks (strong scaling, no collective)."""
    from quantum_b200 import circuits as cq
    from quantum_b200.sharding import row_block
    n = 22
    m, names, qs = cq.tfi_chain_circuit(n)
    prog = cq.serialize(m)
    ham = cq.tfi_hamiltonian(qs)
    G = args.c4_batch
    vals = np.random.default_rng(22).uniform(0, 1, (G, len(names))).astype(np.float32)
    lo, hi = row_block(G, rank, world)
    rows = hi - lo
    down = np.ones((rows, 1), np.float32)
    t0 = time.perf_counter()
    job = ops.DeviceJob("adjoint", [prog] * rows, names, vals[lo:hi], [[ham]] * rows, down,
                        device=local)
    job.run()                      # compiles the specialised reverse passes
    ctx.sync()
    first = time.perf_counter() - t0
    job.run()
    ctx.sync()
    steps = 2
    sec, prof = timed(job.run, steps, profile=True)
    g = job.fetch()
    job.close()
    a_ms = prof["adjoint_pass_ms"]
    ach = prof["adjoint_pass_bytes"] / max(a_ms * 1e-3, 1e-12) / 1e9
    return {"workload": "configs[3]: %d-qubit TFI-chain ansatz, %d symbols, 484 "
                        "parameterised gates, tfq_adj_grad, global batch %d split "
                        "over %d GPU(s)" % (n, len(names), G, world),
            "scaling": "strong", "global_batch": G, "rows_per_gpu": rows,
            "value": G * steps / sec, "unit": "circuits/s",
            "ms_per_step": 1e3 * sec / steps, "steps": steps,
            "first_call_s_max_over_ranks": max_over_ranks(first),
            "grad_abs_mean": float(np.abs(g).mean()) if g.size else None,
            "reverse_pass_roofline": {"bound": "hbm", "achieved": ach, "peak": peak,
                                      "unit": "GB/s", "frac": ach / peak,
                                      "launches": int(prof["adjoint_pass_launches"]),
                                      "share_of_step": a_ms * 1e-3 / sec,
                                      "specialised_launches": int(prof["jit_pass_launches"])}}


def _grid_shape(n):
    r = int(np.floor(np.sqrt(n)))
    while n % r:
        r -= 1
    return r, n // r


def _c5_circuit(n):
    from quantum_b200 import circuits as cq
    rows, cols = _grid_shape(n)
    m, qs = cq.supremacy_style_circuit(rows, cols, 20, n, use_line=True)
    return cq.serialize(m), [cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs])]


def leg_sharded_state(args, ops, ctx, rank, world, local, barrier, max_over_ranks, peak):
    """configs[4]: ONE state, C1-style random circuit, depth 20, sum Z.  On one
    GPU the plain op at 34 qubits (128 GiB); on N = 2, 4, 8 GPUs the state of
    33 + log2(N) qubits is sharded by its top qubits (64 GiB per GPU + the
    exchange buffer) and the global<->local qubit swaps run inside the library
    through peer memory (CUDA IPC over NVLink, quantum_b200/sharded.py)."""
    import torch
    from quantum_b200 import sharded
    g = int(np.log2(world))
    if (1 << g) != world:
        return {"skipped": "world size %d is not a power of two" % world}
    n = args.sharded_qubits or (34 if world == 1 else 33 + g)
    out = {"workload": "configs[4]: single %d-qubit state, random circuit depth 20, "
                       "sum Z, %d GPU(s)" % (n, world),
           "n_qubits": n, "state_gib": 8.0 * 2 ** n / 2 ** 30}
    zeros = np.zeros((1, 0), np.float32)
    try:
        # small-n check of the sharded path against the unsharded op
        if world > 1:
            prog_c, sums_c = _c5_circuit(24)
            a = sharded.peer_sharded_expectation(prog_c, [], zeros[0], sums_c, device=local)
            b = ops.tfq_simulate_expectation([prog_c], [], zeros, [sums_c], device=local)[0]
            out["check"] = {"n_qubits": 24, "sharded": float(a[0]), "unsharded": float(b[0]),
                            "abs_err": float(abs(a[0] - b[0])),
                            "ok": bool(abs(a[0] - b[0]) < 1e-5 + 1e-4 * abs(b[0]))}
            ctx.trim()
        prog, sums = _c5_circuit(n)
        if world == 1:
            def run():
                return ops.tfq_simulate_expectation([prog], [], zeros, [sums], device=local)[0]
            job = None
        else:
            job = sharded.peer_sharded_job(prog, [], zeros[0], sums, device=local)

            def run():
                job.enqueue()
                return job.result()
        run()                                  # interpreted kernels
        run()                                  # compiles the specialised passes
        ctx.profile_reset()
        ctx.profile_enable(True)
        barrier()
        ctx.sync()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e = run()
        ctx.sync()
        sec = max_over_ranks(time.perf_counter() - t0)
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        gp_ms = prof["gate_pass_ms"]
        ach = prof["gate_pass_bytes"] / max(gp_ms * 1e-3, 1e-12) / 1e9
        out.update(seconds_per_circuit=sec, expectation=float(e[0]),
                   gate_passes=int(prof["gate_pass_launches"]),
                   specialised_launches=int(prof["jit_pass_launches"]),
                   gate_pass_roofline={"bound": "hbm", "achieved": ach, "peak": peak,
                                       "unit": "GB/s", "frac": ach / peak,
                                       "avg_pass_ms": gp_ms / max(prof["gate_pass_launches"], 1),
                                       "share_of_circuit": gp_ms * 1e-3 / sec})
        if job is not None:
            st = job.stats()
            ex = st["exchanges"]
            out["exchange"] = {
                "how": "peer memory inside the library (CUDA IPC over NVLink / "
                       "NVSwitch, epoch flags, no host collective)",
                "exchanges": ex, "fused_into_next_gate_pass": st["fused_exchanges"],
                "seconds": st["pull_ms"] * 1e-3,
                "wait_for_peers_seconds": st["wait_ms"] * 1e-3,
                "GBps_received_per_gpu": (st["bytes_received_per_exchange"] * ex /
                                          max(st["pull_ms"] * 1e-3, 1e-12) / 1e9) if ex else None,
                "share_of_circuit": st["pull_ms"] * 1e-3 / sec}
            barrier()          # nobody unmaps a shard a peer may still read
            job.close()
    except Exception as exc:   # noqa: BLE001  (keep the headline line alive)
        out["error"] = "%s: %s" % (type(exc).__name__, exc)
    return out


def leg_one_call_all_gpus(ops, rank, world, barrier):
    """ONE tfq_simulate_expectation call over a context that spans every GPU of
    the node (tfqb_create_multi), made by rank 0 while the other ranks idle:
    what a TF op kernel does when the placer gives it the whole node
    (tfq_simulate_expectation_op.cc:245-248 spreads one Compute over all host
    cores the same way)."""
    barrier()
    out = None
    if rank == 0:
        try:
            rows = BATCH * world
            prog, names, obs, vals, _ = workload(rows, seed=77)
            mctx = ops.Context(list(range(world)))
            ops._contexts[tuple(range(world))] = mctx
            dev = tuple(range(world))
            ops.tfq_simulate_expectation([prog] * rows, names, vals, [obs] * rows, device=dev)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                e = ops.tfq_simulate_expectation([prog] * rows, names, vals, [obs] * rows,
                                                 device=dev)
            sec = (time.perf_counter() - t0) / reps
            # the first device's block through the single-device context
            one = ops.tfq_simulate_expectation([prog] * BATCH, names, vals[:BATCH],
                                               [obs] * BATCH, device=0)
            out = {"workload": "configs[1] circuit, %d rows in ONE call from one process "
                               "over %d GPUs (host buffers in, host buffers out)"
                               % (rows, world),
                   "value": rows / sec, "unit": "circuits/s", "ms_per_call": 1e3 * sec,
                   "devices": mctx.device_count(),
                   "identical_to_single_device": bool(np.array_equal(e[:BATCH], one)),
                   "max_abs_diff_vs_single_device": float(np.abs(e[:BATCH] - one).max())}
            ops._contexts.pop(dev, None)
            mctx.close()
        except Exception as exc:   # noqa: BLE001
            out = {"error": "%s: %s" % (type(exc).__name__, exc)}
    barrier()
    return out


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
    ap.add_argument("--no-extra-legs", action="store_true",
                    help="skip c4_strong / sharded_state / one_call_all_gpus")
    ap.add_argument("--c4-batch", type=int, default=2048,
                    help="GLOBAL batch of the strong-scaling adjoint leg")
    ap.add_argument("--sharded-qubits", type=int, default=0,
                    help="0: 34 on one GPU, 33 + log2(N) on N")
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

    # a wait on the HOST: an NCCL barrier parks a spinning kernel on every
    # waiting rank's GPU, and kernels of another process on that GPU are then
    # time-sliced against it (the one-call leg drives all GPUs from rank 0)
    host_group = dist.new_group(backend="gloo") if world > 1 else None

    def host_barrier():
        if world > 1:
            dist.barrier(group=host_group)

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
    t_cold = time.perf_counter()
    job = ops.DeviceJob("expectation", programs, names, vals, sums, device=local)
    job.run()
    ctx.sync()
    cold = {"first_call_s": time.perf_counter() - t_cold,
            "nvrtc_cpu_s": ops.jit_compile_seconds(),
            "note": "first evaluation of a new circuit structure: parse + plan + "
                    "NVRTC of every pass (parallel host threads; cached on disk "
                    "under TFQB_JIT_CACHE_DIR afterwards) + one step"}
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
        with open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)) as f:
            tj = json.load(f)
        if tj.get("n_qubits") == N_QUBITS:
            n_pass = prof["gate_pass_launches"] / max(K, 1)   # passes per step
            per_state = (tj["forward_first_pass_dram_bytes_per_state"] +
                         (n_pass - 1) * tj["forward_pass_rw_dram_bytes_per_state"]) / n_pass
            traffic = per_state * B
    except Exception:
        traffic = None
    # FP32-pipe floor of the gate passes: packed FP32 instructions per
    # amplitude of every pass (counted by the kernel generator, csrc/jit.cc) x
    # amplitudes / the measured FFMA2 issue rate of this part
    # (profiles/r01f_pipe_rates_b200.jsonl: 70.4 TFLOP/s = 17.6e12 FFMA2/s)
    fp32 = None
    try:
        plan = ops.host_describe_plan(prog, names)
        ppa = [p["packed_fp32_per_amp_phase_free"] for p in plan["passes"]]
        if all(x >= 0 for x in ppa):
            floor_ms = sum(ppa) * 2.0 ** N_QUBITS * B / FFMA2_PER_S * 1e3
            step_gate_ms = gp_ms / K
            fp32 = {"packed_fp32_per_amplitude_per_pass": ppa,
                    "ffma2_rate_per_s": FFMA2_PER_S,
                    "floor_ms_per_step": floor_ms,
                    "gate_pass_ms_per_step": step_gate_ms,
                    "frac": floor_ms / step_gate_ms,
                    "note": "the gate passes are FP32-issue bound at this gate "
                            "density (16 B of HBM traffic buy 67-70 packed "
                            "FP32 instructions per amplitude in passes 0/1): "
                            "frac is the distance to THAT ceiling, roofline.frac "
                            "the distance to the HBM one"}
    except Exception as e:      # noqa: BLE001
        fp32 = {"error": str(e)}
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
                          "profiles/%s (ncu --set full capture of this kernel), "
                          "scaled to this batch" % TRAFFIC_FILE,
        "peak_source": peak_src,
        "launches": int(prof["gate_pass_launches"]),
        "avg_launch_ms": gp_ms / gp_launches,
        "algorithmic_bytes_per_launch": prof["gate_pass_bytes"] / gp_launches,
        "share_of_step": gp_ms * 1e-3 / sec,
        "fp32": fp32,
        "expectation_kernel": {
            "achieved": prof["expectation_bytes"] /
            max(prof["expectation_ms"] * 1e-3, 1e-12) / 1e9,
            "unit": "GB/s", "launches": int(prof["expectation_launches"]),
            "share_of_step": prof["expectation_ms"] * 1e-3 / sec},
    }
    launches = int(prof["kernel_launches"])

    # ---- leg 2: end to end through the public API (host buffers)
    e2e_steps = max(1, min(K, 10))
    # a job that first qualifies hands every pass of its plans to NVRTC on
    # background host threads; the ones nothing has launched yet still compile
    # and would be timed as host contention: part of the warm-up, waited for
    jit_wait_s = ops.wait_for_jit()
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
           "steps": e2e_steps, "waited_for_background_compiles_s": jit_wait_s,
           "note": "ops.tfq_simulate_expectation(host strings + float32 "
                   "arrays) -> numpy; includes proto parse, planning, H2D, D2H"}

    # ---- leg 3: adjoint gradient
    adjoint = None
    if not args.no_adjoint:
        t_cold = time.perf_counter()
        ajob = ops.DeviceJob("adjoint", programs, names, vals, sums, down,
                             device=local)
        ajob.run()
        ctx.sync()
        a_first = time.perf_counter() - t_cold
        for _ in range(W):
            ajob.run()
        ctx.sync()
        asec, aprof = timed(ajob.run, K, profile=True)
        adj_result = ajob.fetch()
        ajob.close()
        a_ms = aprof["adjoint_pass_ms"]
        a_ach = aprof["adjoint_pass_bytes"] / max(a_ms * 1e-3, 1e-12) / 1e9
        ops.wait_for_jit()
        t0 = time.perf_counter()
        ops.tfq_adj_grad(programs, names, vals, sums, down, device=local)
        a_e2e = max_over_ranks(time.perf_counter() - t0)
        adjoint = {
            "metric": "adjoint-grad circuits/sec", "value": world * B * K / asec,
            "unit": UNIT, "ms_per_step": 1e3 * asec / K,
            "e2e": {"value": world * B / a_e2e, "unit": UNIT},
            "first_call_s": a_first,
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
        t, t_prep = cpu_port_time(n_c, cores)
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
                         "%.1f s of C simulation (restated qsim-style CPU path, "
                         "oracle/qsim_vm.c, AVX2 gate loops); the oracle's Python "
                         "preparation (%.1f s) is not charged" % (n_c, t, t_prep)}

    extra = {}
    if not args.no_extra_legs:
        ctx.trim()
        extra["c4_strong"] = leg_c4_strong(args, ops, ctx, rank, world, local, timed,
                                           max_over_ranks, peak)
        ctx.trim()
        extra["sharded_state"] = leg_sharded_state(args, ops, ctx, rank, world, local,
                                                   barrier, max_over_ranks, peak)
        ctx.trim()
        if world > 1:
            torch.cuda.synchronize()
            extra["one_call_all_gpus"] = leg_one_call_all_gpus(ops, rank, world, host_barrier)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * sec / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "complex64", "data": "synthetic",
            "config": config_dict(world, B), "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu, "parity_vs_oracle": parity, "adjoint": adjoint,
            "cold_start": cold,
            "nvrtc_cpu_s_total": ops.jit_compile_seconds(),
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
