#!/usr/bin/env python
"""Two calls of tfq_adj_grad on the C2 workload at a small batch: the launch
pattern ncu captures from (7 tfqb_jit_pass launches per call: 3 forward, 4
reverse) -- scripts/measure_round3.sh."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_b200 import circuits as cq  # noqa: E402
from quantum_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kind = sys.argv[2] if len(sys.argv) > 2 else "adjoint"
m, names, qs = cq.hea_circuit(20, 4)
prog = cq.serialize(m)
obs = cq.hea_observables(qs)
vals = np.random.default_rng(20).uniform(0, 2, (B, len(names))).astype(np.float32)
down = np.ones((B, len(obs)), np.float32)
for _ in range(2):
    if kind == "adjoint":
        g = ops.tfq_adj_grad([prog] * B, names, vals, [obs] * B, down)
    else:
        g = ops.tfq_simulate_expectation([prog] * B, names, vals, [obs] * B)
print(float(np.abs(g).mean()))
