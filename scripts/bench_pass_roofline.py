#!/usr/bin/env python
"""Achieved HBM bandwidth of the gate passes as a function of gate density.

north_star asks for >= 70 % of the B200 HBM bandwidth on fused-gate passes at
20-34 qubits.  A pass moves 16 B per amplitude whatever it applies, so the
fraction depends on how many dense gates the planner packs into it: the FP32
pipe sustains about 11 dense 1-qubit gates per pass at HBM speed (DESIGN.md
section 6).  This script measures it: hardware-efficient circuits of 1..4
layers on n qubits (the first layer is synthesised as a product state, the
planner packs the rest into as few passes as it can), gate-pass time from CUDA
events on the library's stream, bytes = 16 * 2^n * rows per pass (8 for the
write-only first pass).  One JSON line per (n, layers).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_b200 import circuits as cq  # noqa: E402
from quantum_b200 import ops  # noqa: E402


def peak_gbs():
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(
                os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6535.4, "fallback (round-1 measured copy bandwidth)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", default="20,24,28")
    ap.add_argument("--layers", default="1,2,3,4")
    ap.add_argument("--gib", type=float, default=8.0, help="state bytes per launch")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--random-depths", default="6,12,20")
    a = ap.parse_args()
    peak, src = peak_gbs()
    ctx = ops.get_context()
    for n in [int(x) for x in a.qubits.split(",")]:
        rows = max(1, int(a.gib * 2 ** 30 / (8 * 2 ** n)))
        for layers in [int(x) for x in a.layers.split(",")]:
            moments, names, qs = cq.hea_circuit(n, layers)
            prog = cq.serialize(moments)
            d = ops.host_describe_plan(prog, names)
            # every layer after the first (product state) is one fused dense
            # 2x2 gate (Y^a then Z^b) per qubit, plus the CZ sign ops
            n_dense_gates = (layers - 1) * n
            vals = np.random.default_rng(n + layers).uniform(
                0, 2, (rows, len(names))).astype(np.float32)
            obs = [cq.pauli_sum([(1.0, [(qs[0], "Z")])])]
            job = ops.DeviceJob("expectation", [prog] * rows, names, vals, [obs] * rows)
            for _ in range(3):
                job.run()
            ctx.sync()
            ctx.profile_reset()
            ctx.profile_enable(True)
            for _ in range(a.steps):
                job.run()
            ctx.sync()
            prof = ctx.profile_read()
            ctx.profile_enable(False)
            job.close()
            gbs = prof["gate_pass_bytes"] / max(prof["gate_pass_ms"] * 1e-3, 1e-12) / 1e9
            print(json.dumps({
                "circuit": "HEA %d layers" % layers,
                "n_qubits": n, "rows": rows, "layers": layers,
                "passes": len(d["passes"]), "dispatches_per_pass": [p["ops"] for p in d["passes"]],
                "dense_2x2_gates": n_dense_gates,
                "dense_gates_per_pass": round(n_dense_gates / max(len(d["passes"]), 1), 1),
                "gate_pass_ms_per_step": prof["gate_pass_ms"] / a.steps,
                "achieved_GBps": gbs, "peak_GBps": peak, "peak_source": src,
                "frac_of_hbm_peak": gbs / peak,
                "specialised": prof["jit_pass_launches"] > 0}), flush=True)
    random_circuits([int(x) for x in a.qubits.split(",")],
                    [int(x) for x in a.random_depths.split(",")], a.gib, a.steps)


def grid_shape(n):
    r = int(np.floor(np.sqrt(n)))
    while n % r:
        r -= 1
    return r, n // r


def random_circuits(qubits, depths, gib, steps):
    """C1 / C5-style random circuits (CZ layers + X^1/2, Y^1/2, Z^1/4 on idle
    qubits): few dense gates per pass at small depth."""
    peak, src = peak_gbs()
    ctx = ops.get_context()
    for n in qubits:
        rows = max(1, int(gib * 2 ** 30 / (8 * 2 ** n)))
        for depth in depths:
            r, c = grid_shape(n)
            m, qs = cq.supremacy_style_circuit(r, c, depth, n, use_line=True)
            prog = cq.serialize(m)
            d = ops.host_describe_plan(prog, [])
            n_gates = sum(1 for g in d["gates"] if g["kind"] > 1)
            vals = np.zeros((rows, 0), np.float32)
            obs = [cq.pauli_sum([(1.0, [(qs[0], "Z")])])]
            job = ops.DeviceJob("expectation", [prog] * rows, [], vals, [obs] * rows)
            for _ in range(3):
                job.run()
            ctx.sync()
            ctx.profile_reset()
            ctx.profile_enable(True)
            for _ in range(steps):
                job.run()
            ctx.sync()
            prof = ctx.profile_read()
            ctx.profile_enable(False)
            job.close()
            gbs = prof["gate_pass_bytes"] / max(prof["gate_pass_ms"] * 1e-3, 1e-12) / 1e9
            print(json.dumps({
                "circuit": "random depth %d" % depth, "n_qubits": n, "rows": rows,
                "gates": n_gates, "passes": len(d["passes"]),
                "dispatches_per_pass": [p["ops"] for p in d["passes"]],
                "gate_pass_ms_per_step": prof["gate_pass_ms"] / steps,
                "achieved_GBps": gbs, "peak_GBps": peak, "peak_source": src,
                "frac_of_hbm_peak": gbs / peak,
                "specialised": prof["jit_pass_launches"] > 0}), flush=True)


if __name__ == "__main__":
    main()
