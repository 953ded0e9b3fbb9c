#!/usr/bin/env python
"""Where the end-to-end call of the C2 workload spends its host time: the
phases of tfq_simulate_expectation (host buffers in, host buffers out) timed
separately through the device-resident job API (prepare / run / fetch)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from quantum_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prog, names, obs, vals, _ = bench.workload(B)
programs, sums = [prog] * B, [obs] * B
ctx = ops.get_context()
for _ in range(3):
    ops.tfq_simulate_expectation(programs, names, vals, sums)
out = {"rows": B}
reps = 5
t = {"marshal": 0.0, "prepare": 0.0, "run": 0.0, "fetch": 0.0, "free": 0.0, "one_call": 0.0}
for _ in range(reps):
    t0 = time.perf_counter()
    inp = ops._Inputs(programs, names, vals)
    pk = ops._pauli_pack(sums)
    t1 = time.perf_counter()
    job = ops.DeviceJob("expectation", programs, names, vals, sums)
    t2 = time.perf_counter()
    job.run()
    ctx.sync()
    t3 = time.perf_counter()
    job.fetch()
    t4 = time.perf_counter()
    job.close()
    t5 = time.perf_counter()
    ops.tfq_simulate_expectation(programs, names, vals, sums)
    t6 = time.perf_counter()
    t["marshal"] += t1 - t0
    t["prepare"] += (t2 - t1) - (t1 - t0)      # DeviceJob marshals again
    t["run"] += t3 - t2
    t["fetch"] += t4 - t3
    t["free"] += t5 - t4
    t["one_call"] += t6 - t5
out.update({k + "_ms": 1e3 * v / reps for k, v in t.items()})
print(json.dumps(out))
