#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs[0], [2], [3], [4]) through the
public API (quantum_b200.ops); one JSON line per config.  These are parity /
sanity workloads, not the headline bench line (bench.py)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_b200 import circuits as cq  # noqa: E402
from quantum_b200 import ops  # noqa: E402


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), out


def c1(batch=100):
    m, qs = cq.supremacy_style_circuit(2, 5, 20, 63536323)
    prog = cq.serialize(m)
    ps = cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs])
    vals = np.zeros((batch, 0), np.float32)
    t, _ = timed(lambda: ops.tfq_simulate_expectation([prog] * batch, [], vals,
                                                      [[ps]] * batch))
    return {"config": "C1 10q random depth 20, batch %d, Expectation Z-sum" % batch,
            "circuits_per_s": batch / t, "ms": 1e3 * t}


def c3(batch=32, n=24, shots=1000):
    qs = [cq.grid(0, i) for i in range(n)]
    progs = [cq.serialize(cq.random_circuit(qs, 20, 24 + r)) for r in range(batch)]
    vals = np.zeros((batch, 0), np.float32)
    t, out = timed(lambda: ops.tfq_simulate_samples(progs, [], vals, [shots], seed=7),
                   reps=2)
    ps = cq.pauli_sum([(1.0, [(qs[i], "Z"), (qs[i + 1], "Z")]) for i in range(n - 1)] +
                      [(1.0, [(q, "X")]) for q in qs])
    ns = np.full((batch, 1), shots, np.int32)
    t2, _ = timed(lambda: ops.tfq_simulate_sampled_expectation(
        progs, [], vals, [[ps]] * batch, ns, seed=7), reps=1)
    return {"config": "C3 %dq random circuits (distinct per row), batch %d, %d shots"
                      % (n, batch, shots),
            "samples_circuits_per_s": batch / t, "samples_ms": 1e3 * t,
            "sampled_expectation_circuits_per_s": batch / t2,
            "sampled_expectation_ms": 1e3 * t2, "ones_fraction": float((out == 1).mean())}


def c4(batch=256, n=22):
    m, names, qs = cq.tfi_chain_circuit(n)
    prog = cq.serialize(m)
    ham = cq.tfi_hamiltonian(qs)
    vals = np.random.default_rng(22).uniform(0, 1, (batch, len(names))).astype(np.float32)
    down = np.ones((batch, 1), np.float32)
    t, g = timed(lambda: ops.tfq_adj_grad([prog] * batch, names, vals,
                                          [[ham]] * batch, down), reps=2)
    t2, _ = timed(lambda: ops.tfq_simulate_expectation([prog] * batch, names, vals,
                                                       [[ham]] * batch), reps=2)
    return {"config": "C4 %dq TFI-chain VQE ansatz, %d symbols, batch %d"
                      % (n, len(names), batch),
            "adjoint_circuits_per_s": batch / t, "adjoint_ms": 1e3 * t,
            "expectation_circuits_per_s": batch / t2, "grad_norm": float(np.abs(g).mean())}


def c5(n=30):
    rows = 5
    m, qs = cq.supremacy_style_circuit(rows, n // rows, 20, n, use_line=True)
    prog = cq.serialize(m)
    ps = cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs])
    vals = np.zeros((1, 0), np.float32)
    t, e = timed(lambda: ops.tfq_simulate_expectation([prog], [], vals, [[ps]]), reps=3)
    return {"config": "C5 single %dq state, depth 20, Z-sum" % len(qs),
            "seconds_per_circuit": t, "expectation": float(e[0, 0])}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c1,c3,c4,c5")
    ap.add_argument("--c5-qubits", type=int, default=30)
    ap.add_argument("--c4-batch", type=int, default=256)
    ap.add_argument("--c3-batch", type=int, default=32)
    a = ap.parse_args()
    for name in a.only.split(","):
        fn = {"c1": c1, "c3": lambda: c3(a.c3_batch), "c4": lambda: c4(a.c4_batch),
              "c5": lambda: c5(a.c5_qubits)}[name]
        print(json.dumps(fn()), flush=True)
