"""Cirq-free circuit builder + serializer to TFQ's wire format, and the
synthetic workloads of BASELINE.json.

A circuit is a list of moments; a moment is a list of `Op`.  `serialize()`
emits the same bytes TFQ's `serializer.py` would (gate ids and arg names per
tensorflow_quantum/core/serialize/serializer.py:456-486,527-560,645-683;
qubit ids per op_serializer.py:33-39), so the strings can be fed to the real
TFQ ops unchanged.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np

from . import proto

Param = Union[float, str]  # literal value or symbol name

EIGEN_1Q = ("XP", "YP", "ZP", "HP")
EIGEN_2Q = ("XXP", "YYP", "ZZP", "CZP", "CNP", "SP", "ISP")


def grid(row: int, col: int) -> str:
    """GridQubit id string (op_serializer.py:33-39)."""
    return f"{row}_{col}"


def line(x: int) -> str:
    """LineQubit id string."""
    return f"{x}"


@dataclass
class Op:
    gate: str
    qubits: Tuple[str, ...]
    args: Dict[str, Param] = field(default_factory=dict)
    scalars: Dict[str, float] = field(default_factory=dict)
    controls: Tuple[str, ...] = ()
    control_values: Tuple[int, ...] = ()

    def controlled_by(self, controls, values=None) -> "Op":
        controls = tuple(controls)
        if values is None:
            values = (1,) * len(controls)
        return Op(self.gate, self.qubits, dict(self.args), dict(self.scalars),
                  self.controls + controls,
                  self.control_values + tuple(values))


def _r6(x):
    return float(np.round(float(x), 6))


def eigen(gate: str, qubits: Sequence[str], exponent: Param = 1.0,
          scalar: float = 1.0, global_shift: float = 0.0) -> Op:
    assert gate in EIGEN_1Q + EIGEN_2Q
    return Op(gate, tuple(qubits), {"exponent": exponent,
                                    "global_shift": float(global_shift)},
              {"exponent": scalar})


def X(q, exponent=1.0, scalar=1.0, global_shift=0.0):
    return eigen("XP", (q,), exponent, scalar, global_shift)


def Y(q, exponent=1.0, scalar=1.0, global_shift=0.0):
    return eigen("YP", (q,), exponent, scalar, global_shift)


def Z(q, exponent=1.0, scalar=1.0, global_shift=0.0):
    return eigen("ZP", (q,), exponent, scalar, global_shift)


def H(q, exponent=1.0, scalar=1.0, global_shift=0.0):
    return eigen("HP", (q,), exponent, scalar, global_shift)


def rx(q, rads: Param, scalar=1.0):
    """cirq.rx: XPowGate(exponent=rads/pi, global_shift=-0.5)."""
    if isinstance(rads, str):
        return eigen("XP", (q,), rads, scalar / np.pi, -0.5)
    return eigen("XP", (q,), rads / np.pi, 1.0, -0.5)


def XX(a, b, exponent=1.0, scalar=1.0, global_shift=0.0):
    return eigen("XXP", (a, b), exponent, scalar, global_shift)


def YY(a, b, exponent=1.0, scalar=1.0, global_shift=0.0):
    return eigen("YYP", (a, b), exponent, scalar, global_shift)


def ZZ(a, b, exponent=1.0, scalar=1.0, global_shift=0.0):
    return eigen("ZZP", (a, b), exponent, scalar, global_shift)


def CZ(a, b, exponent=1.0, scalar=1.0, global_shift=0.0):
    return eigen("CZP", (a, b), exponent, scalar, global_shift)


def CNOT(a, b, exponent=1.0, scalar=1.0, global_shift=0.0):
    return eigen("CNP", (a, b), exponent, scalar, global_shift)


def SWAP(a, b, exponent=1.0, scalar=1.0, global_shift=0.0):
    return eigen("SP", (a, b), exponent, scalar, global_shift)


def ISWAP(a, b, exponent=1.0, scalar=1.0, global_shift=0.0):
    return eigen("ISP", (a, b), exponent, scalar, global_shift)


def I(q):
    return Op("I", (q,))


def I2(a, b):
    return Op("I2", (a, b))


def PhasedX(q, phase_exponent: Param, exponent: Param = 1.0,
            phase_scalar=1.0, scalar=1.0, global_shift=0.0):
    return Op("PXP", (q,), {"phase_exponent": phase_exponent,
                            "exponent": exponent,
                            "global_shift": float(global_shift)},
              {"phase_exponent": phase_scalar, "exponent": scalar})


def FSim(a, b, theta: Param, phi: Param, theta_scalar=1.0, phi_scalar=1.0):
    return Op("FSIM", (a, b), {"theta": theta, "phi": phi},
              {"theta": theta_scalar, "phi": phi_scalar})


def PhasedISwap(a, b, phase_exponent: Param, exponent: Param = 1.0,
                phase_scalar=1.0, scalar=1.0):
    return Op("PISP", (a, b), {"phase_exponent": phase_exponent,
                               "exponent": exponent, "global_shift": 0.0},
              {"phase_exponent": phase_scalar, "exponent": scalar})


# ---- noise channels (serializer.py:175-450: ids DP ADP GAD AD RST PD PF BF;
# literal float args only -- "cirq channels can't contain symbols") ---------
def depolarize(q, p):
    return Op("DP", (q,), {"p": float(p)})


def asymmetric_depolarize(q, p_x, p_y, p_z):
    return Op("ADP", (q,), {"p_x": float(p_x), "p_y": float(p_y), "p_z": float(p_z)})


def generalized_amplitude_damp(q, p, gamma):
    return Op("GAD", (q,), {"p": float(p), "gamma": float(gamma)})


def amplitude_damp(q, gamma):
    return Op("AD", (q,), {"gamma": float(gamma)})


def reset(q):
    return Op("RST", (q,))


def phase_damp(q, gamma):
    return Op("PD", (q,), {"gamma": float(gamma)})


def phase_flip(q, p):
    return Op("PF", (q,), {"p": float(p)})


def bit_flip(q, p):
    return Op("BF", (q,), {"p": float(p)})


def _set_param(op_pb, name, val):
    if isinstance(val, str):
        op_pb.args[name].symbol = val
    else:
        op_pb.args[name].arg_value.float_value = _r6(val)


def to_program(moments: Sequence[Sequence[Op]]) -> "proto.Program":
    p = proto.Program()
    p.language.gate_set = "tfq_gate_set"
    p.circuit.scheduling_strategy = proto.MOMENT_BY_MOMENT
    for moment in moments:
        m = p.circuit.moments.add()
        for op in moment:
            o = m.operations.add()
            o.gate.id = op.gate
            for name, val in op.args.items():
                _set_param(o, name, val)
            for name, s in op.scalars.items():
                o.args[name + "_scalar"].arg_value.float_value = _r6(s)
            if op.gate == "I":
                o.args["unused"].arg_value.bool_values.values.append(True)
            o.args["control_qubits"].arg_value.string_value = ",".join(
                op.controls)
            o.args["control_values"].arg_value.string_value = ",".join(
                str(int(v)) for v in op.control_values)
            for q in op.qubits:
                o.qubits.add().id = q
    return p


def serialize(moments: Sequence[Sequence[Op]]) -> bytes:
    """Circuit -> bytes, as `tfq.convert_to_tensor` would produce
    (tensorflow_quantum/python/util.py:264-336)."""
    return to_program(moments).SerializeToString()


def serialize_text(moments) -> bytes:
    """Text-format variant (what benchmark_random_circuit.py:103 feeds)."""
    from google.protobuf import text_format
    return text_format.MessageToString(to_program(moments)).encode()


PauliTermSpec = Tuple[float, Sequence[Tuple[str, str]]]


def pauli_sum(terms: Sequence[PauliTermSpec]) -> bytes:
    """[(coeff, [(qubit_id, 'X'|'Y'|'Z'), ...]), ...] -> PauliSum bytes."""
    ps = proto.PauliSum()
    for coeff, paulis in terms:
        t = ps.terms.add()
        t.coefficient_real = float(coeff)
        t.coefficient_imag = 0.0
        for qid, ptype in paulis:
            pq = t.paulis.add()
            pq.qubit_id = qid
            pq.pauli_type = ptype
    return ps.SerializeToString()


# --------------------------------------------------------------------------
# Workloads (SURVEY.md §8d table "Concrete synthetic inputs")
# --------------------------------------------------------------------------

def hea_circuit(n: int, layers: int):
    """C2: hardware-efficient ansatz. Per layer Y^s then Z^s on every qubit
    (fresh symbol each), then CZ on even pairs then odd pairs."""
    qs = [grid(0, i) for i in range(n)]
    moments, names = [], []
    for l in range(layers):
        ym, zm = [], []
        for i, q in enumerate(qs):
            ys, zs = f"y{l}_{i}", f"z{l}_{i}"
            names += [ys, zs]
            ym.append(Y(q, ys))
            zm.append(Z(q, zs))
        moments += [ym, zm]
        moments.append([CZ(qs[i], qs[i + 1]) for i in range(0, n - 1, 2)])
        if n > 2:
            moments.append([CZ(qs[i], qs[i + 1]) for i in range(1, n - 1, 2)])
    return moments, names, qs


def hea_observables(qs):
    """C2: M=4 sums: sum Z; sum ZZ; sum X; sum 0.5 XY + 0.25 I."""
    n = len(qs)
    return [
        pauli_sum([(1.0, [(q, "Z")]) for q in qs]),
        pauli_sum([(1.0, [(qs[i], "Z"), (qs[i + 1], "Z")])
                   for i in range(n - 1)]),
        pauli_sum([(1.0, [(q, "X")]) for q in qs]),
        pauli_sum([(0.5, [(qs[i], "X"), (qs[i + 1], "Y")])
                   for i in range(n - 1)] + [(0.25, [])]),
    ]


def tfi_chain_circuit(n: int, depth: int = None):
    """C4: TFI-chain VQE ansatz (reference datasets/spin_system.py:254-261):
    H on all; per layer ZZ^{s_d} on the ring bonds, X^{s_{d+depth}} on all."""
    if depth is None:
        depth = n // 2
    qs = [grid(0, i) for i in range(n)]
    names = [f"theta_{i}" for i in range(2 * depth)]
    moments = [[H(q) for q in qs]]
    for d in range(depth):
        zz = names[d]
        x = names[d + depth]
        moments.append([ZZ(qs[i], qs[i + 1], zz) for i in range(0, n - 1, 2)])
        odd = [ZZ(qs[i], qs[i + 1], zz) for i in range(1, n - 1, 2)]
        if n > 2 and n % 2 == 0:
            odd.append(ZZ(qs[n - 1], qs[0], zz))
            moments.append(odd)
        else:
            moments.append(odd)
            if n > 2:
                moments.append([ZZ(qs[n - 1], qs[0], zz)])
        moments.append([X(q, x) for q in qs])
    return moments, names, qs


def tfi_hamiltonian(qs, g: float = 1.0):
    """-sum ZZ (ring) - g sum X (reference spin_system.py:302-306)."""
    n = len(qs)
    terms = [(-1.0, [(qs[i], "Z"), (qs[(i + 1) % n], "Z")])
             for i in range(n if n > 2 else n - 1)]
    terms += [(-g, [(q, "X")]) for q in qs]
    return pauli_sum(terms)


_CZ_PATTERNS = 8


def supremacy_style_circuit(rows: int, cols: int, depth: int, seed: int,
                            use_line=False):
    """C1/C5: restated 'supremacy-style' random circuit: moment 0 = H on all;
    middle moments = a CZ layer from an 8-pattern cycle plus one of
    {X^1/2, Y^1/2, Z^1/4} on idle qubits; last moment = H on all
    (reference: benchmarks/scripts/benchmark_random_circuit.py:36-42 uses
    cirq's generate_boixo_2018_supremacy_circuits_v2_grid; cirq is absent, so
    the distribution is restated, not bit-identical)."""
    rng = np.random.default_rng(seed)
    qid = (lambda r, c: line(r * cols + c)) if use_line else grid
    qs = [qid(r, c) for r in range(rows) for c in range(cols)]
    moments = [[H(q) for q in qs]]
    for d in range(depth - 2):
        pat = d % _CZ_PATTERNS
        horizontal = pat % 2 == 0
        shift = (pat // 2) % 2
        stagger = (pat // 4) % 2
        busy, m = set(), []
        if horizontal:
            for r in range(rows):
                start = (shift + stagger * (r % 2)) % 2
                for c in range(start, cols - 1, 2):
                    m.append(CZ(qid(r, c), qid(r, c + 1)))
                    busy |= {(r, c), (r, c + 1)}
        else:
            for c in range(cols):
                start = (shift + stagger * (c % 2)) % 2
                for r in range(start, rows - 1, 2):
                    m.append(CZ(qid(r, c), qid(r + 1, c)))
                    busy |= {(r, c), (r + 1, c)}
        for r in range(rows):
            for c in range(cols):
                if (r, c) in busy:
                    continue
                k = int(rng.integers(3))
                q = qid(r, c)
                m.append([X(q, 0.5), Y(q, 0.5), Z(q, 0.25)][k])
        moments.append(m)
    moments.append([H(q) for q in qs])
    return moments, qs


_RANDOM_1Q = ("XP", "YP", "ZP", "HP", "PXP", "I")
_RANDOM_2Q = ("XXP", "YYP", "ZZP", "CZP", "CNP", "SP", "ISP", "FSIM", "PISP",
              "I2")


def _random_gate(rng, gate, qubits, exponent: Param, scalar=1.0):
    if gate in EIGEN_1Q + EIGEN_2Q:
        return eigen(gate, qubits, exponent, scalar)
    if gate == "PXP":
        return PhasedX(qubits[0], 0.123, exponent, 1.0, scalar)
    if gate == "PISP":
        return PhasedISwap(qubits[0], qubits[1], 0.123, exponent, 1.0, scalar)
    if gate == "FSIM":
        if isinstance(exponent, str):
            return FSim(qubits[0], qubits[1], exponent, 0.456, scalar, 1.0)
        return FSim(qubits[0], qubits[1], 0.123 * exponent, 0.456 * exponent)
    if gate == "I":
        return I(qubits[0])
    return I2(qubits[0], qubits[1])


def random_circuit(qs: Sequence[str], n_moments: int, seed: int, p: float = 0.9,
                   controls: bool = False, symbols: Sequence[str] = (),
                   include_scalars: bool = True):
    """Random circuit over the TFQ gate set with random exponents U(0,1)
    (distribution of reference python/util.py:124-214; optional random
    controls as python/util.py:94-121). If `symbols` is given, one gate per
    moment takes a (scaled) symbol instead of a number."""
    rng = np.random.default_rng(seed)
    qs = list(qs)
    moments = []
    sym_i = 0
    for mi in range(n_moments):
        free = list(qs)
        rng.shuffle(free)
        m = []
        sym_done = not symbols
        while free:
            q = free.pop()
            if rng.random() > p:
                continue
            two = len(free) > 0 and rng.random() < 0.4
            if two:
                q2 = free.pop()
                gate = _RANDOM_2Q[int(rng.integers(len(_RANDOM_2Q)))]
                targets = (q, q2)
            else:
                gate = _RANDOM_1Q[int(rng.integers(len(_RANDOM_1Q)))]
                targets = (q,)
            if not sym_done and gate not in ("I", "I2"):
                expo = symbols[sym_i % len(symbols)]
                sym_i += 1
                sym_done = True
                scalar = float(np.round(rng.random(), 6)) \
                    if include_scalars else 1.0
            else:
                expo = float(rng.random())
                scalar = 1.0
            op = _random_gate(rng, gate, targets, expo, scalar)
            if controls and rng.random() < 0.5:
                open_q = [x for x in qs if x not in targets]
                k = min(len(open_q), 3)
                if k > 0:
                    k = int(rng.integers(1, k + 1))
                    idx = rng.choice(len(open_q), size=k, replace=False)
                    cq = [open_q[int(i)] for i in idx]
                    cv = [int(v) for v in rng.integers(0, 2, size=k)]
                    op = op.controlled_by(cq, cv)
                    # controls occupy those qubits in this moment
                    free = [x for x in free if x not in cq]
            m.append(op)
        if m:
            moments.append(m)
    # Use the rest of the symbols (python/util.py:166-170)
    while symbols and sym_i < len(symbols):
        moments.append([H(qs[0], symbols[sym_i])])
        sym_i += 1
    return moments


def random_pauli_sum(qs: Sequence[str], max_terms: int, seed: int,
                     max_weight: int = None) -> bytes:
    """reference python/util.py:241-257 (random_pauli_sums)."""
    rng = np.random.default_rng(seed)
    n_terms = int(rng.integers(1, max_terms + 1))
    terms = []
    mw = len(qs) if max_weight is None else min(max_weight, len(qs))
    for _ in range(n_terms):
        w = int(rng.integers(1, mw + 1))
        idx = rng.choice(len(qs), size=w, replace=False)
        paulis = []
        for i in idx:
            p = "IXYZ"[int(rng.integers(4))]
            if p != "I":
                paulis.append((qs[int(i)], p))
        terms.append((float(np.round(rng.normal(), 4)), paulis))
    return pauli_sum(terms)
