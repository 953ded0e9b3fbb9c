#!/usr/bin/env python
"""Yardstick for the gradient parity tests at BASELINE sizes: the reference's
adjoint algorithm with its float32 gate matrices but the STATE carried in
complex128 (oracle backend "numpy128").  The distance of the float32 oracle
from this yardstick is the float32 round-off floor of the algorithm itself —
what qsim's complex64 simulation cannot do better than — and the GPU path is
held to the same distance (tests/test_gpu_parity_at_size.py).

Writes tests/golden/f64_state_gradients.npz.  CPU only; the 22-qubit case
takes about ten minutes of numpy.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tfq_oracle as orc          # noqa: E402
from quantum_b200 import circuits as cq        # noqa: E402


def c2():
    moments, names, qs = cq.hea_circuit(20, 4)
    prog = cq.serialize(moments)
    obs = cq.hea_observables(qs)
    vals = np.random.default_rng(20).uniform(0, 2, (2, len(names))).astype(np.float32)
    down = np.ones((2, 4), np.float32)
    return [prog] * 2, names, vals, [obs] * 2, down


def c4():
    m, names, qs = cq.tfi_chain_circuit(22)
    prog = cq.serialize(m)
    ham = cq.tfi_hamiltonian(qs)
    vals = np.random.default_rng(22).uniform(0, 1, (1, len(names))).astype(np.float32)
    return [prog], names, vals, [[ham]], np.ones((1, 1), np.float32)


if __name__ == "__main__":
    out = {}
    orc.INNER_THREADS = os.cpu_count() or 1
    for name, mk in (("c2", c2), ("c4", c4)):
        progs, names, vals, sums, down = mk()
        t0 = time.time()
        g64 = orc.adjoint_gradient(progs, names, vals, sums, down, backend="numpy128",
                                   out_dtype=np.float64)
        g32 = orc.adjoint_gradient(progs, names, vals, sums, down)
        out[name + "_f64_state"] = g64
        out[name + "_oracle_f32"] = g32
        print(name, "%.0f s" % (time.time() - t0), "oracle(float32) vs double state: max abs",
              float(np.abs(g32 - g64).max()), "scale", float(np.abs(g64).max()), flush=True)
    np.savez(os.path.join(ROOT, "tests", "golden", "f64_state_gradients.npz"), **out)
