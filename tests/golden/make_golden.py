#!/usr/bin/env python
"""Generates tests/golden/*.npz: seeded inputs of the five ops together with
the CPU oracle's outputs (oracle/tfq_oracle.py, itself pinned to the
reference's golden vectors in tests/test_oracle_goldens.py).  The real
reference cannot be imported here (no TensorFlow / cirq / qsim in the image),
so these fixtures freeze the oracle; GPU parity tests compare against them so
that a later change of the oracle cannot silently move the target.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import tfq_oracle as orc  # noqa: E402
from quantum_b200 import circuits as cq  # noqa: E402


def obj(x):
    a = np.empty(len(x), dtype=object)
    for i, v in enumerate(x):
        a[i] = v
    return a


def main():
    names = ["a", "b", "c"]
    n_list = [2, 3, 5, 7, 9, 12]
    progs, sums = [], []
    for k, n in enumerate(n_list):
        qs = [cq.grid(0, i) for i in range(n)] if k % 2 == 0 else \
            [cq.grid(1, i) for i in range(n // 2)] + [cq.line(i) for i in range(n - n // 2)]
        progs.append(cq.serialize(cq.random_circuit(qs, 8, 1000 + k, controls=True,
                                                    symbols=names)))
        sums.append([cq.random_pauli_sum(qs, 5, 50 + 7 * k + j, max_weight=4)
                     for j in range(2)])
    progs.append(cq.serialize([]))
    sums.append([cq.pauli_sum([(1.0, [])])] * 2)
    B = len(progs)
    rng = np.random.default_rng(2024)
    vals = rng.uniform(0, 2, (B, 3)).astype(np.float32)
    down = rng.normal(size=(B, 2)).astype(np.float32)
    S = 64
    u = rng.random((B, S))
    ns = rng.integers(1, S + 1, size=(B, 2)).astype(np.int32)
    ue = rng.random((B, 2, 5, S))
    np.savez_compressed(
        os.path.join(HERE, "five_ops_ragged.npz"),
        programs=obj(progs), symbol_names=np.array(names),
        symbol_values=vals, pauli_sums=obj([obj(r) for r in sums]),
        downstream=down, uniforms=u, num_samples=ns, uniforms_exp=ue,
        state=orc.simulate_state(progs, names, vals),
        expectation=orc.simulate_expectation(progs, names, vals, sums),
        gradient=orc.adjoint_gradient(progs, names, vals, sums, down),
        samples=orc.simulate_samples(progs, names, vals, [S], uniforms=u),
        sampled_expectation=orc.simulate_sampled_expectation(
            progs, names, vals, sums, ns, uniforms=ue))
    # BASELINE configs[1] shape at a size the oracle finishes in seconds
    m, hn, qs = cq.hea_circuit(12, 4)
    prog = cq.serialize(m)
    obs = cq.hea_observables(qs)
    v = np.random.default_rng(20).uniform(0, 2, (5, len(hn))).astype(np.float32)
    np.savez_compressed(
        os.path.join(HERE, "hea12_expectation_adjoint.npz"),
        program=np.array([prog], dtype=object), symbol_names=np.array(hn),
        symbol_values=v, pauli_sums=obj(obs),
        expectation=orc.simulate_expectation([prog] * 5, hn, v, [obs] * 5),
        gradient=orc.adjoint_gradient([prog] * 5, hn, v, [obs] * 5,
                                      np.ones((5, 4), np.float32)))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
