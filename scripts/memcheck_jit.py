#!/usr/bin/env python
"""Small workload that drives every run-time specialised kernel (forward exact
and phase-free, expectation, accumulate, adjoint) plus the interpreted
fallbacks; meant to run under `compute-sanitizer --tool memcheck` with
TFQB_JIT_MIN_AMPS=0."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TFQB_JIT_MIN_AMPS", "0")
from quantum_b200 import circuits as cq  # noqa: E402
from quantum_b200 import ops  # noqa: E402

ctx = ops.get_context()
rng = np.random.default_rng(0)
for name, (mo, names, qs), obs in (
        ("hea", cq.hea_circuit(13, 3), None),
        ("tfi", cq.tfi_chain_circuit(13, 2), "tfi")):
    prog = cq.serialize(mo)
    sums = cq.hea_observables(qs) if obs is None else [cq.tfi_hamiltonian(qs)]
    v = rng.uniform(0, 2, (3, len(names))).astype(np.float32)
    ops.tfq_simulate_state([prog] * 3, names, v)
    ops.tfq_simulate_expectation([prog] * 3, names, v, [sums] * 3)
    ops.tfq_adj_grad([prog] * 3, names, v, [sums] * 3, np.ones((3, len(sums)), np.float32))
    ops.tfq_simulate_samples([prog] * 3, names, v, [16])
qs = [cq.grid(0, i) for i in range(13)]
for seed, controls in ((1, False), (2, True)):
    m = cq.random_circuit(qs, 10, seed, controls=controls, symbols=("a", "b"))
    prog = cq.serialize(m)
    v = rng.uniform(0, 2, (2, 2)).astype(np.float32)
    sums = [cq.random_pauli_sum(qs, 6, 3, max_weight=4)]
    ops.tfq_simulate_expectation([prog] * 2, ["a", "b"], v, [sums] * 2)
    ops.tfq_adj_grad([prog] * 2, ["a", "b"], v, [sums] * 2, np.ones((2, 1), np.float32))
    oth = cq.serialize(cq.random_circuit(qs, 4, 9) + [[cq.H(q) for q in qs]])
    ops.tfq_inner_product_grad([prog] * 2, ["a", "b"], v, [[oth]] * 2, np.ones((2, 1), np.float32))
print("done", ctx.profile_read()["jit_kernels"], "specialised kernels compiled,",
      ctx.profile_read()["jit_pass_launches"], "specialised launches")
