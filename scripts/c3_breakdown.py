#!/usr/bin/env python
"""Where configs[2] (24-qubit DISTINCT random circuits, 1000 shots) spends its
time: gate passes (CUDA events) against wall clock, per circuit."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_b200 import circuits as cq  # noqa: E402
from quantum_b200 import ops  # noqa: E402

B, n, S = 64, 24, 1000
qs = [cq.grid(0, i) for i in range(n)]
progs = [cq.serialize(cq.random_circuit(qs, 20, 24 + r)) for r in range(B)]
vals = np.zeros((B, 0), np.float32)
ctx = ops.get_context()
ops.tfq_simulate_samples(progs[:4], [], vals[:4], [S], seed=1)          # warm-up
out = {}
for label, timing in (("wall_no_events", False), ("with_events", True)):
    ctx.profile_reset()
    ctx.profile_enable(timing)
    t0 = time.perf_counter()
    ops.tfq_simulate_samples(progs, [], vals, [S], seed=7)
    wall = time.perf_counter() - t0
    p = ctx.profile_read()
    ctx.profile_enable(False)
    out[label] = {"ms_per_circuit": 1e3 * wall / B,
                  "gate_pass_ms_per_circuit": p["gate_pass_ms"] / B,
                  "gate_passes_per_circuit": p["gate_pass_launches"] / B,
                  "gate_pass_GBps": p["gate_pass_bytes"] / max(p["gate_pass_ms"], 1e-9) / 1e6,
                  "kernel_launches_per_circuit": p["kernel_launches"] / B}
# the same programs a second time: parse / lower / plan are cached by program bytes
t0 = time.perf_counter()
ops.tfq_simulate_samples(progs, [], vals, [S], seed=7)
out["second_call_ms_per_circuit"] = 1e3 * (time.perf_counter() - t0) / B
t0 = time.perf_counter()
ops.tfq_simulate_state(progs[:8], [], vals[:8])
out["state_ms_per_circuit_incl_128MiB_D2H"] = 1e3 * (time.perf_counter() - t0) / 8
print(json.dumps(out))
