"""CPU ORACLE — test infrastructure only, never shipped, never on the product path.

A restatement of TensorFlow Quantum's qsim CPU path for the five
circuit-execution ops.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import this.

Parity status: PINNED against the reference's own golden vectors
(tests/test_oracle_goldens.py): util_qsim_test.cc:49-181,236-355,437-528,
adj_util_test.cc:355-560, tfq_adj_grad_op_test.py:265-397,
circuit_execution_ops.py:52-68, gradient_test.py:230-248.
UNPINNED by any offline golden (checked by the reference only against a live
cirq): closed forms of HP/ZP/ZZP/YYP/CZP/SP/ISP (restated from Cirq's public
definitions; unitarity + identities tested) and sample bitstrings (the
"identical uniforms" contract is defined here, see `sample_tree`).

The arithmetic itself lives in qsim v0.21.0 (WORKSPACE:74-83, sha256
720eeb97...9617), which is NOT under /root/reference; its published algorithm
is restated: float32 gate matrices from Cirq closed forms, complex64 state,
fp64 reductions, gate-by-gate sweeps with <=2-qubit basic fusion.

Layout: (1) proto parse + qubit resolution, (2) gate builders, (3) circuit /
gradient-circuit construction, (4) lowering of each op to a small program of
state-space steps (the reference's orchestration, sweep for sweep), (5) two
interchangeable executors for those steps: numpy (`NumpyVM`) and the C
restatement in qsim_vm.c (`CVM`, used for the timed CPU baseline).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from quantum_b200 import proto as _pb

F32 = np.float32
GRAD_EPS = F32(5e-3)  # adj_util.cc:32


class InvalidArgumentError(ValueError):
    """Stands in for tf.errors.InvalidArgumentError."""


# ==========================================================================
# (1) parsing and qubit-id resolution
# ==========================================================================

def parse_proto(data: bytes, cls):
    """parse_context.cc:41-56: binary first, then text format."""
    if isinstance(data, str):
        data = data.encode()
    msg = cls()
    try:
        msg.ParseFromString(data)
        ok = True
    except Exception:
        ok = False
    if ok:
        return msg
    try:
        from google.protobuf import text_format
        msg = cls()
        text_format.Parse(data.decode(), msg)
        return msg
    except Exception:
        raise InvalidArgumentError("Unparseable proto: " + repr(data[:40]))


_INT_MAX = 2147483647


def _register_qubits(qb_string: str, id_set: set):
    """program_resolution.cc:49-88."""
    if qb_string == "":
        return
    for qb in qb_string.split(","):
        splits = qb.split("_")
        if len(splits) == 1:
            splits = [str(_INT_MAX)] + splits
        if len(splits) != 2:
            raise InvalidArgumentError("Unable to parse qubit: " + qb)
        try:
            r, c = int(splits[0]), int(splits[1])
        except ValueError:
            raise InvalidArgumentError("Unable to parse qubit: " + qb)
        id_set.add(((r, c), qb))


def resolve_qubit_ids(program, p_sums=None) -> int:
    """program_resolution.cc:90-186. Rewrites ids in place; returns n."""
    if len(program.circuit.moments) == 0:
        return 0
    id_set = set()
    for moment in program.circuit.moments:
        for op in moment.operations:
            for q in op.qubits:
                _register_qubits(q.id, id_set)
            if "control_qubits" not in op.args:
                raise InvalidArgumentError("missing control_qubits arg")
            _register_qubits(op.args["control_qubits"].arg_value.string_value,
                             id_set)
    ids = sorted(id_set)
    id_to_index = {s: str(i) for i, (_, s) in enumerate(ids)}
    for moment in program.circuit.moments:
        for op in moment.operations:
            for q in op.qubits:
                q.id = id_to_index[q.id]
            cq = op.args["control_qubits"].arg_value.string_value
            if cq == "":
                continue
            op.args["control_qubits"].arg_value.string_value = ",".join(
                id_to_index[c] for c in cq.split(","))
    if p_sums is not None:
        for ps in p_sums:
            for term in ps.terms:
                for pair in term.paulis:
                    if pair.qubit_id not in id_to_index:
                        raise InvalidArgumentError(
                            "Found a Pauli sum operating on qubits not found "
                            "in circuit.")
                    pair.qubit_id = id_to_index[pair.qubit_id]
    return len(ids)


# ==========================================================================
# (2) gate matrices — Cirq closed forms in float32.
# Convention: a 2-qubit matrix is written over the operation's own qubit
# order (a, b): row/col index = 2*x_a + x_b.
#
# Canonical float32 recipe (shared bit-for-bit with csrc/gates.cuh so that
# finite-difference gradient gates are common-mode):
#   ang  = f32(pi32 * t)                      (float32 product)
#   c,s  = f32(cos/sin(f64(ang) * 0.5))       (double trig, rounded once)
#   g    = f32(cos/sin(f64(ang) * (0.5 + f64(shift))))
#   entries = single float32 products / sums of those, no FMA contraction.
# ==========================================================================

PI32 = F32(3.14159265358979323846)


def _f(x):
    return F32(x)


def _cs(arg64):
    return F32(np.cos(arg64)), F32(np.sin(arg64))


def _cmul(ar, ai, br, bi):
    """(ar + i ai)(br + i bi) in float32, products rounded individually."""
    ar, ai, br, bi = F32(ar), F32(ai), F32(br), F32(bi)
    return F32(F32(ar * br) - F32(ai * bi)), F32(F32(ar * bi) + F32(ai * br))


def _cplx(re, im):
    return np.complex64(complex(float(re), float(im)))


def _eigen_parts(t, shift):
    """Returns (c, s, gr, gi, g0r, g0i): half-angle cos/sin, the phase
    g = e^{i pi t (0.5+shift)} and g0 = e^{i pi t shift}."""
    t = F32(t)
    shift = F32(shift)
    ang = np.float64(F32(PI32 * t))
    c, s = _cs(ang * 0.5)
    gr, gi = _cs(ang * (0.5 + np.float64(shift)))
    g0r, g0i = _cs(ang * np.float64(shift))
    return c, s, gr, gi, g0r, g0i


def mat_xpow(t, shift=0.0):
    c, s, gr, gi, _, _ = _eigen_parts(t, shift)
    d = _cplx(F32(c * gr), F32(c * gi))           # c g
    o = _cplx(F32(s * gi), F32(-F32(s * gr)))     # -i s g
    return np.array([[d, o], [o, d]], dtype=np.complex64)


def mat_ypow(t, shift=0.0):
    c, s, gr, gi, _, _ = _eigen_parts(t, shift)
    d = _cplx(F32(c * gr), F32(c * gi))
    p = _cplx(F32(s * gr), F32(s * gi))           # s g
    return np.array([[d, -p], [p, d]], dtype=np.complex64)


def mat_zpow(t, shift=0.0):
    t = F32(t)
    shift = F32(shift)
    ang = np.float64(F32(PI32 * t))
    g0r, g0i = _cs(ang * np.float64(shift))
    g1r, g1i = _cs(ang * (1.0 + np.float64(shift)))
    return np.array([[_cplx(g0r, g0i), 0], [0, _cplx(g1r, g1i)]],
                    dtype=np.complex64)


IS2 = F32(0.70710678118654752440)


def mat_hpow(t, shift=0.0):
    c, s, gr, gi, _, _ = _eigen_parts(t, shift)
    # g (c I - i s H),  H = (X+Z)/sqrt2
    ar = F32(F32(s * gi) * IS2)            # Re(-i s g)/sqrt2
    ai = F32(-F32(F32(s * gr) * IS2))      # Im(-i s g)/sqrt2
    dr, di = F32(c * gr), F32(c * gi)
    return np.array([[_cplx(F32(dr + ar), F32(di + ai)), _cplx(ar, ai)],
                     [_cplx(ar, ai), _cplx(F32(dr - ar), F32(di - ai))]],
                    dtype=np.complex64)


def mat_xxpow(t, shift=0.0):
    c, s, gr, gi, _, _ = _eigen_parts(t, shift)
    d = _cplx(F32(c * gr), F32(c * gi))
    o = _cplx(F32(s * gi), F32(-F32(s * gr)))     # -i s g
    m = np.zeros((4, 4), dtype=np.complex64)
    for i in range(4):
        m[i, i] = d
        m[i, 3 - i] = o
    return m


def mat_yypow(t, shift=0.0):
    c, s, gr, gi, _, _ = _eigen_parts(t, shift)
    d = _cplx(F32(c * gr), F32(c * gi))
    o = _cplx(F32(s * gi), F32(-F32(s * gr)))     # -i s g
    m = np.zeros((4, 4), dtype=np.complex64)
    for i in range(4):
        m[i, i] = d
    m[0, 3] = -o
    m[3, 0] = -o
    m[1, 2] = o
    m[2, 1] = o
    return m


def mat_zzpow(t, shift=0.0):
    z = mat_zpow(t, shift)
    return np.diag([z[0, 0], z[1, 1], z[1, 1], z[0, 0]]).astype(np.complex64)


def mat_czpow(t, shift=0.0):
    z = mat_zpow(t, shift)
    return np.diag([z[0, 0], z[0, 0], z[0, 0], z[1, 1]]).astype(np.complex64)


def mat_cxpow(t, shift=0.0):
    x = mat_xpow(t, shift)
    z = mat_zpow(t, shift)
    m = np.zeros((4, 4), dtype=np.complex64)
    m[0, 0] = m[1, 1] = z[0, 0]
    m[2:, 2:] = x
    return m


def mat_swappow(t, shift=0.0):
    x = mat_xpow(t, shift)
    z = mat_zpow(t, shift)
    m = np.zeros((4, 4), dtype=np.complex64)
    m[0, 0] = m[3, 3] = z[0, 0]
    m[1, 1] = m[2, 2] = x[0, 0]
    m[1, 2] = m[2, 1] = x[0, 1]
    return m


def mat_iswappow(t, shift=0.0):
    t = F32(t)
    shift = F32(shift)
    ang = np.float64(F32(PI32 * t))
    c, s = _cs(ang * 0.5)
    g0r, g0i = _cs(ang * np.float64(shift))
    dr, di = F32(c * g0r), F32(c * g0i)               # c g0
    orr, oi = F32(-F32(s * g0i)), F32(s * g0r)        # i s g0
    m = np.zeros((4, 4), dtype=np.complex64)
    m[0, 0] = m[3, 3] = _cplx(g0r, g0i)
    m[1, 1] = m[2, 2] = _cplx(dr, di)
    m[1, 2] = m[2, 1] = _cplx(orr, oi)
    return m


def mat_phasedxpow(p, t, shift=0.0):
    c, s, gr, gi, _, _ = _eigen_parts(t, shift)
    pang = np.float64(F32(PI32 * F32(p)))
    pr, pi_ = _cs(pang)                               # e^{i pi p}
    d = _cplx(F32(c * gr), F32(c * gi))
    o_r, o_i = F32(s * gi), F32(-F32(s * gr))         # -i s g
    ur, ui = _cmul(o_r, o_i, pr, F32(-pi_))           # * e^{-i pi p}
    lr, li = _cmul(o_r, o_i, pr, pi_)                 # * e^{+i pi p}
    return np.array([[d, _cplx(ur, ui)], [_cplx(lr, li), d]],
                    dtype=np.complex64)


def mat_fsim(theta, phi):
    theta = np.float64(F32(theta))
    phi = np.float64(F32(phi))
    c, s = _cs(theta)
    pr, pi_ = _cs(phi)
    m = np.zeros((4, 4), dtype=np.complex64)
    m[0, 0] = 1
    m[1, 1] = m[2, 2] = _cplx(c, 0)
    m[1, 2] = m[2, 1] = _cplx(0, F32(-s))
    m[3, 3] = _cplx(pr, F32(-pi_))
    return m


def mat_phasediswappow(p, t):
    ang = np.float64(F32(PI32 * F32(t)))
    c, s = _cs(ang * 0.5)
    fang = np.float64(F32(PI32 * F32(p))) * 2.0
    fr, fi = _cs(fang)                                # f = e^{2 pi i p}
    m = np.zeros((4, 4), dtype=np.complex64)
    m[0, 0] = m[3, 3] = 1
    m[1, 1] = m[2, 2] = _cplx(c, 0)
    m[1, 2] = _cplx(F32(-F32(s * fi)), F32(s * fr))   # i s f
    m[2, 1] = _cplx(F32(s * fi), F32(s * fr))         # i s conj(f)
    return m


EIGEN_1Q = {"XP": mat_xpow, "YP": mat_ypow, "ZP": mat_zpow, "HP": mat_hpow}
EIGEN_2Q = {"XXP": mat_xxpow, "YYP": mat_yypow, "ZZP": mat_zzpow,
            "CZP": mat_czpow, "CNP": mat_cxpow, "SP": mat_swappow,
            "ISP": mat_iswappow}


# ==========================================================================
# (3) circuit construction (circuit_parser_qsim.cc:53-596,828-861) and
#     gradient circuit (adj_util.cc:37-302)
# ==========================================================================

@dataclass
class Gate:
    kind: str
    qubits: Tuple[int, ...]          # resolved proto indices == axes
    matrix: np.ndarray               # complex64 [2^k, 2^k], op qubit order
    controls: Tuple[int, ...] = ()
    cvalues: Tuple[int, ...] = ()
    # adjoint metadata (circuit_parser_qsim.h:35-65)
    params: Tuple[float, ...] = ()
    symbols: List[str] = field(default_factory=list)
    placeholders: List[str] = field(default_factory=list)


def _arg(op, name, smap, used=None):
    """ParseProtoArg, circuit_parser_qsim.cc:53-82."""
    if name not in op.args:
        raise InvalidArgumentError("Could not find arg: " + name + " in op.")
    a = op.args[name]
    val = F32(a.arg_value.float_value)
    if a.symbol != "":
        if a.symbol not in smap:
            raise InvalidArgumentError(
                "Could not find symbol in parameter map: " + a.symbol)
        val = F32(smap[a.symbol][1])
        if used is not None:
            used.append(a.symbol)
    return val


def _controls(op):
    """ParseProtoControls, circuit_parser_qsim.cc:84-129."""
    cs = op.args["control_qubits"].arg_value.string_value
    cv = op.args["control_values"].arg_value.string_value
    if cs == "" and cv == "":
        return (), ()
    ct, vt = cs.split(","), cv.split(",")
    if len(ct) != len(vt):
        raise InvalidArgumentError(
            "Mistmatched number of control qubits and control values.")
    try:
        vals = tuple(int(v) for v in vt)
    except ValueError:
        raise InvalidArgumentError("Unparseable control value: " + cv)
    return tuple(int(c) for c in ct), vals


def build_gate(op, smap) -> Gate:
    gid = op.gate.id
    qs = tuple(int(q.id) for q in op.qubits)
    ctr, cv = _controls(op) if gid in _ALL_IDS else ((), ())
    if gid in EIGEN_1Q or gid in EIGEN_2Q:
        used = []
        e = _arg(op, "exponent", smap, used)
        es = _arg(op, "exponent_scalar", smap)
        gs = _arg(op, "global_shift", smap)
        fn = EIGEN_1Q.get(gid) or EIGEN_2Q.get(gid)
        g = Gate(gid, qs, fn(F32(e * es), gs), ctr, cv, (e, es, gs))
        if used:
            g.symbols, g.placeholders = used, ["exponent"]
        return g
    if gid == "PXP":
        ue, up = [], []
        e = _arg(op, "exponent", smap, ue)
        es = _arg(op, "exponent_scalar", smap)
        pe = _arg(op, "phase_exponent", smap, up)
        pes = _arg(op, "phase_exponent_scalar", smap)
        gs = _arg(op, "global_shift", smap)
        g = Gate(gid, qs, mat_phasedxpow(F32(pe * pes), F32(e * es), gs), ctr,
                 cv, (pe, pes, e, es, gs))
        if up:
            g.symbols.append(up[0]); g.placeholders.append("phase_exponent")
        if ue:
            g.symbols.append(ue[0]); g.placeholders.append("exponent")
        return g
    if gid == "FSIM":
        ut, uph = [], []
        th = _arg(op, "theta", smap, ut)
        ths = _arg(op, "theta_scalar", smap)
        ph = _arg(op, "phi", smap, uph)
        phs = _arg(op, "phi_scalar", smap)
        g = Gate(gid, qs, mat_fsim(F32(th * ths), F32(ph * phs)), ctr, cv,
                 (th, ths, ph, phs))
        if ut:
            g.symbols.append(ut[0]); g.placeholders.append("theta")
        if uph:
            g.symbols.append(uph[0]); g.placeholders.append("phi")
        return g
    if gid == "PISP":
        ue, up = [], []
        e = _arg(op, "exponent", smap, ue)
        es = _arg(op, "exponent_scalar", smap)
        pe = _arg(op, "phase_exponent", smap, up)
        pes = _arg(op, "phase_exponent_scalar", smap)
        g = Gate(gid, qs, mat_phasediswappow(F32(pe * pes), F32(e * es)), ctr,
                 cv, (pe, pes, e, es))
        if up:
            g.symbols.append(up[0]); g.placeholders.append("phase_exponent")
        if ue:
            g.symbols.append(ue[0]); g.placeholders.append("exponent")
        return g
    if gid == "I":
        return Gate(gid, qs, np.eye(2, dtype=np.complex64), ctr, cv)
    if gid == "I2":
        return Gate(gid, qs, np.eye(4, dtype=np.complex64), ctr, cv)
    raise InvalidArgumentError(
        "Could not parse gate id: " + gid + ". This is likely because a "
        "cirq.Channel was used in an op that does not support them.")


_ALL_IDS = set(EIGEN_1Q) | set(EIGEN_2Q) | {"PXP", "FSIM", "PISP", "I", "I2"}


def circuit_from_program(program, smap, n) -> List[Gate]:
    """QsimCircuitFromProgram (circuit_parser_qsim.cc:828-861), gate list in
    moment order.  Returns [] for n <= 0."""
    if n <= 0:
        return []
    gates = []
    for moment in program.circuit.moments:
        for op in moment.operations:
            gates.append(build_gate(op, smap))
    return gates


def _rebuild(g: Gate, which: str, delta) -> np.ndarray:
    """Gate matrix with the *unscaled* symbol value shifted by delta
    (adj_util.cc:175-302: `(exp + eps) * exp_s`)."""
    d = F32(delta)
    if g.kind in EIGEN_1Q or g.kind in EIGEN_2Q:
        e, es, gs = g.params
        fn = EIGEN_1Q.get(g.kind) or EIGEN_2Q.get(g.kind)
        return fn(F32(F32(e + d) * es), gs)
    if g.kind == "PXP":
        pe, pes, e, es, gs = g.params
        if which == "phase_exponent":
            return mat_phasedxpow(F32(F32(pe + d) * pes), F32(e * es), gs)
        return mat_phasedxpow(F32(pe * pes), F32(F32(e + d) * es), gs)
    if g.kind == "FSIM":
        th, ths, ph, phs = g.params
        if which == "theta":
            return mat_fsim(F32(F32(th + d) * ths), F32(ph * phs))
        return mat_fsim(F32(th * ths), F32(F32(ph + d) * phs))
    if g.kind == "PISP":
        pe, pes, e, es = g.params
        if which == "phase_exponent":
            return mat_phasediswappow(F32(F32(pe + d) * pes), F32(e * es))
        return mat_phasediswappow(F32(pe * pes), F32(F32(e + d) * es))
    raise AssertionError(g.kind)


def _c64_f32(op, a, b):
    """Component-wise float32 op on complex64 arrays."""
    re = op(a.real.astype(F32), b.real.astype(F32)).astype(F32)
    im = op(a.imag.astype(F32), b.imag.astype(F32)).astype(F32)
    return (re + 1j * im).astype(np.complex64)


def gradient_matrix(g: Gate, which: str) -> np.ndarray:
    """(G(p+eps) - G(p-eps)) * (0.5/eps) in float32 (adj_util.h:104-117,
    adj_util.cc:182-187)."""
    left = _rebuild(g, which, GRAD_EPS)
    right = _rebuild(g, which, -GRAD_EPS)
    diff = _c64_f32(np.subtract, left, right)
    scale = F32(0.5 / float(GRAD_EPS))   # 0.5 / _GRAD_EPS: double -> float arg
    re = (diff.real.astype(F32) * scale).astype(F32)
    im = (diff.imag.astype(F32) * scale).astype(F32)
    return (re + 1j * im).astype(np.complex64)


@dataclass
class GradGate:
    index: int                       # position in the gate list
    symbols: List[str]
    matrices: List[np.ndarray]


def gradient_gates(gates: List[Gate]) -> List[GradGate]:
    """CreateGradientCircuit, adj_util.cc:37-153."""
    out = []
    for i, g in enumerate(gates):
        if not g.symbols:
            continue
        out.append(GradGate(i, list(g.symbols),
                            [gradient_matrix(g, w) for w in g.placeholders]))
    return out


# ---- basic <=2-qubit fusion (restating qsim::BasicGateFuser's effect: the
# sweep count of the reference path; call sites circuit_parser_qsim.cc:857,
# adj_util.cc:155-172). A fused gate is (qubits, matrix, controls, cvalues).

def _embed_1q(m1, pos):
    """1q matrix acting on slot `pos` (0 = first/most-significant) of a pair."""
    eye = np.eye(2, dtype=np.complex64)
    return (np.kron(m1, eye) if pos == 0 else np.kron(eye, m1)).astype(
        np.complex64)


_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]],
                 dtype=np.complex64)


def basic_fuse(gates: Sequence[Gate]) -> List[Gate]:
    open_of: Dict[int, int] = {}      # qubit -> index into clusters
    clusters: List[Optional[Gate]] = []
    order: List[int] = []

    def close(q):
        ci = open_of.pop(q, None)
        if ci is None:
            return
        for qq in clusters[ci].qubits:
            if open_of.get(qq) == ci:
                del open_of[qq]
        order.append(ci)

    for g in gates:
        if g.controls:
            for q in g.qubits + g.controls:
                close(q)
            clusters.append(g)
            order.append(len(clusters) - 1)
            continue
        if len(g.qubits) == 1:
            q = g.qubits[0]
            ci = open_of.get(q)
            if ci is None:
                clusters.append(Gate("fused", (q,), g.matrix.copy()))
                open_of[q] = len(clusters) - 1
            else:
                c = clusters[ci]
                m = g.matrix if len(c.qubits) == 1 else _embed_1q(
                    g.matrix, c.qubits.index(q))
                c.matrix = (m @ c.matrix).astype(np.complex64)
            continue
        a, b = g.qubits
        ca, cb = open_of.get(a), open_of.get(b)
        if ca is not None and ca == cb and len(clusters[ca].qubits) == 2:
            c = clusters[ca]
            m = g.matrix if c.qubits == (a, b) else (
                _SWAP @ g.matrix @ _SWAP).astype(np.complex64)
            c.matrix = (m @ c.matrix).astype(np.complex64)
            continue
        m = g.matrix.copy()
        for pos, q in enumerate((a, b)):
            ci = open_of.get(q)
            if ci is not None and len(clusters[ci].qubits) == 1:
                m = (m @ _embed_1q(clusters[ci].matrix, pos)).astype(
                    np.complex64)
                clusters[ci] = None           # absorbed
                del open_of[q]
            else:
                close(q)
        clusters.append(Gate("fused", (a, b), m))
        open_of[a] = open_of[b] = len(clusters) - 1
    for q in sorted(open_of):
        if q in open_of:
            close(q)
    return [clusters[i] for i in order if clusters[i] is not None]


# ==========================================================================
# (4) state-space step programs + executors
# ==========================================================================
# Step tuples (buffers are small ints 0..2 = sv, scratch, scratch2):
#   ("zero_state", buf)                 SetStateZero
#   ("zeros", buf)                      SetAllZeros
#   ("copy", src, dst)
#   ("scale", coeff, buf)               Multiply
#   ("add", src, dst)                   Add
#   ("apply", buf, axes, matrix, caxes, cvals)      ApplyGate / controlled
#   ("project", buf, caxes, cvals)      BulkSetAmpl(.., 0,0, exclude=true)
#   ("dot", a, b, coeff, slot)          out[slot] += coeff*RealInnerProduct

def dagger(m):
    return np.ascontiguousarray(m.conj().T)


class NumpyVM:
    """`dtype=np.complex128` (backend "numpy128") keeps the reference's
    float32 gate matrices but carries the STATE in double: the yardstick that
    separates float32 state round-off (which the reference has too) from
    errors of an implementation (scripts/float32_floor.py)."""

    def __init__(self, n, dtype=np.complex64):
        self.n = max(n, 1)
        self.dtype = dtype
        self.bufs = [np.zeros(2 ** self.n, dtype=dtype)
                     for _ in range(3)]

    def run(self, steps, n_out):
        out = np.zeros(n_out, dtype=np.float64)
        n = self.n
        B = self.bufs
        for st in steps:
            k = st[0]
            if k == "zero_state":
                B[st[1]][:] = 0
                B[st[1]][0] = 1
            elif k == "zeros":
                B[st[1]][:] = 0
            elif k == "copy":
                B[st[2]][:] = B[st[1]]
            elif k == "scale":
                c = F32(st[1])
                b = B[st[2]]
                b.real[:] = b.real * c
                b.imag[:] = b.imag * c
            elif k == "add":
                if self.dtype == np.complex128:
                    B[st[2]][:] = B[st[2]] + B[st[1]]
                else:
                    B[st[2]][:] = _c64_f32(np.add, B[st[2]], B[st[1]])
            elif k == "apply":
                _np_apply(B[st[1]], n, st[2], st[3], st[4], st[5], self.dtype)
            elif k == "project":
                psi = B[st[1]].reshape((2,) * n)
                keep = np.zeros((2,) * n, dtype=bool)
                idx = [slice(None)] * n
                for a, v in zip(st[2], st[3]):
                    idx[a] = v
                keep[tuple(idx)] = True
                psi[~keep] = 0
            elif k == "dot":
                a, b = B[st[1]], B[st[2]]
                v = np.dot(a.real.astype(np.float64), b.real.astype(np.float64)) \
                    + np.dot(a.imag.astype(np.float64),
                             b.imag.astype(np.float64))
                out[st[4]] += F32(st[3]) * v if st[3] is not None else v
            else:
                raise AssertionError(k)
        return out

    def state(self, buf=0):
        return self.bufs[buf]


def _np_apply(state, n, axes, m, caxes=(), cvals=(), dtype=np.complex64):
    k = len(axes)
    psi = state.reshape((2,) * n)
    idx = [slice(None)] * n
    for a, v in zip(caxes, cvals):
        idx[a] = v
    sub = psi[tuple(idx)]
    rem = [a for a in range(n) if a not in caxes]
    pos = [rem.index(a) for a in axes]
    moved = np.moveaxis(sub, pos, list(range(k)))
    shp = moved.shape
    res = (m.astype(dtype) @ moved.reshape(2 ** k, -1)).astype(
        dtype).reshape(shp)
    psi[tuple(idx)] = np.moveaxis(res, list(range(k)), pos)


# ---- C executor (qsim_vm.c) ----------------------------------------------

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force=False):
    """Compile oracle/qsim_vm.c -> oracle/_build/libqsim_vm.so (gcc)."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libqsim_vm.so")
    src = os.path.join(_HERE, "qsim_vm.c")
    if force or not os.path.exists(so) or \
            os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp",
             "-fPIC", "-shared", "-pthread", "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
        _LIB.qvm_run_batch.restype = ctypes.c_int
        _LIB.qvm_run_batch2.restype = ctypes.c_int
    return _LIB


_OPC = {"zero_state": 0, "zeros": 1, "copy": 2, "scale": 3, "add": 4,
        "apply": 5, "project": 6, "dot": 7}


def encode_steps(steps, n):
    """Flatten a step list to (int64 code[], float32 data[]) for qsim_vm.c.
    Axes are converted to bit positions (bit = n-1-axis)."""
    code, data = [], []
    for st in steps:
        k = st[0]
        if k in ("zero_state", "zeros"):
            code += [_OPC[k], st[1]]
        elif k in ("copy", "add"):
            code += [_OPC[k], st[1], st[2]]
        elif k == "scale":
            code += [_OPC[k], st[2], len(data)]
            data.append(float(st[1]))
        elif k == "apply":
            axes, m, caxes, cvals = st[2], st[3], st[4], st[5]
            cmask = sum(1 << (n - 1 - a) for a in caxes)
            cbits = sum((v & 1) << (n - 1 - a) for a, v in zip(caxes, cvals))
            code += [_OPC[k], st[1], len(axes)] + \
                [n - 1 - a for a in axes] + [cmask, cbits, len(data)]
            mm = np.asarray(m, dtype=np.complex64).reshape(-1)
            data += list(mm.view(np.float32))
        elif k == "project":
            cmask = sum(1 << (n - 1 - a) for a in st[2])
            cbits = sum((v & 1) << (n - 1 - a) for a, v in zip(st[2], st[3]))
            code += [_OPC[k], st[1], cmask, cbits]
        elif k == "dot":
            code += [_OPC[k], st[1], st[2], st[4], len(data),
                     0 if st[3] is None else 1]
            data.append(0.0 if st[3] is None else float(st[3]))
    code.append(-1)
    return np.asarray(code, dtype=np.int64), np.asarray(data, dtype=np.float32)


# Timing of the last run_batch_c call: `encode_s` (Python: step lists ->
# code arrays) and `run_s` (the C VM alone).  bench.py's CPU legs report
# `run_s` as the simulation time and the Python preparation separately.
LAST_TIMING = {"encode_s": 0.0, "run_s": 0.0}
# OpenMP threads inside each sweep of a circuit (ComputeLarge,
# tfq_simulate_expectation_op.cc:130-180); 1 = one thread per circuit only.
INNER_THREADS = 1


def run_batch_c(programs, n_list, n_out_list, threads=1, want_state=False):
    """Run a batch of step programs on the C VM, `threads` circuits at a time
    (one thread per circuit: ComputeSmall, tfq_simulate_expectation_op.cc:
    182-250), INNER_THREADS threads inside each sweep (ComputeLarge).
    Returns list of fp64 output arrays (and final sv if asked)."""
    import time as _time
    lib = _lib()
    nb = len(programs)
    _t0 = _time.perf_counter()
    enc = [encode_steps(p, max(n, 1)) for p, n in zip(programs, n_list)]
    codes = (ctypes.c_void_p * nb)(*[e[0].ctypes.data for e in enc])
    datas = (ctypes.c_void_p * nb)(*[e[1].ctypes.data for e in enc])
    ns = np.asarray([max(n, 1) for n in n_list], dtype=np.int32)
    outs = [np.zeros(max(k, 1), dtype=np.float64) for k in n_out_list]
    outp = (ctypes.c_void_p * nb)(*[o.ctypes.data for o in outs])
    states = None
    statep = None
    if want_state:
        states = [np.zeros(2 ** max(n, 1), dtype=np.complex64) for n in n_list]
        statep = (ctypes.c_void_p * nb)(*[s.ctypes.data for s in states])
    _t1 = _time.perf_counter()
    rc = lib.qvm_run_batch2(ctypes.c_int(nb), ns.ctypes.data_as(ctypes.c_void_p),
                            codes, datas, outp, statep, ctypes.c_int(threads),
                            ctypes.c_int(max(1, int(INNER_THREADS))))
    LAST_TIMING["encode_s"] = _t1 - _t0
    LAST_TIMING["run_s"] = _time.perf_counter() - _t1
    if rc != 0:
        raise RuntimeError("qvm_run_batch failed: %d" % rc)
    return (outs, states) if want_state else outs


def _run(steps, n, n_out, backend, want_state=False):
    if backend in ("numpy", "numpy128"):
        vm = NumpyVM(n, np.complex128 if backend == "numpy128" else np.complex64)
        out = vm.run(steps, max(n_out, 1))
        return (out, vm.state(0).copy()) if want_state else out
    res = run_batch_c([steps], [n], [n_out], 1, want_state)
    if want_state:
        return res[0][0], res[1][0]
    return res[0]


# ==========================================================================
# lowering of the reference orchestration to steps
# ==========================================================================

SV, SCRATCH, SCRATCH2 = 0, 1, 2


def _apply_step(buf, g: Gate, m=None):
    return ("apply", buf, g.qubits, g.matrix if m is None else m,
            g.controls, g.cvalues)


def forward_steps(gates, fuse=True):
    fused = basic_fuse(gates) if fuse else list(gates)
    return [("zero_state", SV)] + [_apply_step(SV, g) for g in fused]


_PAULI = {
    "X": np.array([[0, 1], [1, 0]], dtype=np.complex64),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex64),
    "Z": np.array([[1, 0], [0, -1]], dtype=np.complex64),
}


def _pauli_term_gates(term) -> List[Gate]:
    """QsimCircuitFromPauliTerm (circuit_parser_qsim.cc:863-895): XP/YP/ZP
    with exponent 1, shift 0."""
    fn = {"X": mat_xpow, "Y": mat_ypow, "Z": mat_zpow}
    return [Gate(p.pauli_type + "P", (int(p.qubit_id),),
                 fn[p.pauli_type](1.0, 0.0)) for p in term.paulis]


def _zbasis_gates(term) -> List[Gate]:
    """QsimZBasisCircuitFromPauliTerm (circuit_parser_qsim.cc:897-945):
    X -> Y^-0.5, Y -> X^+0.5, Z -> nothing."""
    out = []
    for p in term.paulis:
        if p.pauli_type == "Z":
            continue
        if p.pauli_type == "Y":
            out.append(Gate("XP", (int(p.qubit_id),), mat_xpow(0.5, 0.0)))
        else:
            out.append(Gate("YP", (int(p.qubit_id),), mat_ypow(-0.5, 0.0)))
    return out


def expectation_steps(p_sum, slot):
    """ComputeExpectationQsim, util_qsim.h:142-188. Identity terms are
    returned separately (they add coefficient_real directly)."""
    steps, ident = [], 0.0
    for term in p_sum.terms:
        if len(term.paulis) == 0:
            ident = float(F32(F32(ident) + F32(term.coefficient_real)))
            steps.append(("ident", slot, F32(term.coefficient_real)))
            continue
        steps.append(("copy", SV, SCRATCH))
        for g in basic_fuse(_pauli_term_gates(term)):
            steps.append(_apply_step(SCRATCH, g))
        steps.append(("dot", SV, SCRATCH, F32(term.coefficient_real), slot))
    return steps


def accumulate_steps(p_sums, coeffs, source, scratch, dest):
    """AccumulateOperators, util_qsim.h:362-414."""
    steps = [("copy", source, scratch), ("zeros", dest)]
    for ps, oc in zip(p_sums, coeffs):
        for term in ps.terms:
            lead = F32(F32(oc) * F32(term.coefficient_real))
            if abs(float(lead)) < 1e-5:
                continue
            if len(term.paulis) != 0:
                for g in basic_fuse(_pauli_term_gates(term)):
                    steps.append(_apply_step(scratch, g))
            steps += [("scale", lead, scratch), ("add", scratch, dest),
                      ("copy", source, scratch)]
    return steps


def adjoint_steps(gates, grads: List[GradGate], p_sums, down, sym_col):
    """TfqAdjointGradientOp::ComputeSmall, tfq_adj_grad_op.cc:199-277.
    Output slot = symbol column."""
    steps = forward_steps(gates)
    steps += accumulate_steps(p_sums, down, SV, SCRATCH2, SCRATCH)
    bounds = [gg.index for gg in grads]
    segs, left = [], 0
    for b in bounds:
        segs.append(basic_fuse(gates[left:b]))
        left = b + 1
    segs.append(basic_fuse(gates[left:]))
    for j in range(len(segs) - 1, -1, -1):
        for f in reversed(segs[j]):
            md = dagger(f.matrix)
            steps.append(_apply_step(SV, f, md))
            steps.append(_apply_step(SCRATCH, f, md))
        if j == 0:
            break
        gg = grads[j - 1]
        cur = gates[gg.index]
        cd = dagger(cur.matrix)
        steps.append(_apply_step(SV, cur, cd))
        for sym, dm in zip(gg.symbols, gg.matrices):
            steps.append(("copy", SV, SCRATCH2))
            if cur.controls:
                steps.append(("project", SCRATCH2, cur.controls, cur.cvalues))
            steps.append(("apply", SCRATCH2, cur.qubits, dm, (), ()))
            col = sym_col[sym]
            steps.append(("dot", SCRATCH2, SCRATCH, None, col))
            steps.append(("dot", SCRATCH, SCRATCH2, None, col))
        steps.append(_apply_step(SCRATCH, cur, cd))
    return steps


def _strip_ident(steps):
    """Split ('ident', slot, c) pseudo-steps out of a step list."""
    real, idents = [], []
    for s in steps:
        (idents if s[0] == "ident" else real).append(s)
    return real, idents


# ==========================================================================
# sampling (our contract; see module docstring)
# ==========================================================================

def prob_tree(state: np.ndarray) -> List[np.ndarray]:
    """Canonical fp64 pairwise tree over p_i = re^2 + im^2 (level 0 = leaves).
    The summation order is fixed by the tree, so any parallel implementation
    that forms the same pairs reproduces every node bit-for-bit."""
    re = state.real.astype(np.float64)
    im = state.imag.astype(np.float64)
    t = re * re + im * im
    levels = [t]
    while len(t) > 1:
        t = t[0::2] + t[1::2]
        levels.append(t)
    return levels


def sample_tree(state: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """Index for each uniform u in [0,1): r = u * norm, then descend the tree:
    go left if r < T[left] else r -= T[left] and go right.  In exact
    arithmetic this is qsim's 'first k with r < csum_k' walk (StateSpace::
    Sample, call site tfq_simulate_samples_op.cc:170)."""
    levels = prob_tree(state)
    u = np.asarray(uniforms, dtype=np.float64)
    r = u * levels[-1][0]
    j = np.zeros(u.shape, dtype=np.int64)
    for lvl in range(len(levels) - 2, -1, -1):
        left = levels[lvl][2 * j]
        go_right = ~(r < left)
        r = np.where(go_right, r - left, r)
        j = 2 * j + go_right.astype(np.int64)
    return j


# Philox4x32-10, counter = (shot, row, stream_a, stream_b), key = seed
_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox_uniforms(seed: int, row: int, stream_a: int, stream_b: int,
                    count: int) -> np.ndarray:
    """u = ((x0 << 32 | x1) >> 11) * 2^-53 from Philox4x32-10."""
    c0 = np.arange(count, dtype=np.uint32)
    c1 = np.full(count, row, dtype=np.uint32)
    c2 = np.full(count, stream_a, dtype=np.uint32)
    c3 = np.full(count, stream_b, dtype=np.uint32)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _PH_M0
            p1 = c2.astype(np.uint64) * _PH_M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _PH_W0)
            k1 = np.uint32(k1 + _PH_W1)
    x = (c0.astype(np.uint64) << np.uint64(32)) | c1.astype(np.uint64)
    return (x >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


# ==========================================================================
# (5) the five ops (semantics: SURVEY.md Appendix D)
# ==========================================================================

def _symbol_maps(symbol_names, symbol_values):
    """GetSymbolMaps, parse_context.cc:291-347."""
    names = [s.decode() if isinstance(s, bytes) else s for s in symbol_names]
    vals = np.asarray(symbol_values, dtype=np.float32)
    if vals.ndim != 2:
        raise InvalidArgumentError("symbol_values must be rank 2.")
    if len(names) != vals.shape[1]:
        raise InvalidArgumentError(
            "Input symbol names and value sizes do not match.")
    return [{nm: (j, vals[i, j]) for j, nm in enumerate(names)}
            for i in range(vals.shape[0])]


def _prologue(programs, symbol_names, symbol_values, pauli_sums=None):
    if np.ndim(programs) != 1:
        raise InvalidArgumentError("programs must be rank 1. Got rank %d."
                                   % np.ndim(programs))
    if np.ndim(symbol_names) != 1:
        raise InvalidArgumentError("symbol_names must be rank 1.")
    progs = [parse_proto(p, _pb.Program) for p in programs]
    sums = None
    if pauli_sums is not None:
        if np.ndim(np.empty((0, 0)) if len(pauli_sums) == 0 else
                   np.asarray(pauli_sums, dtype=object)) != 2:
            raise InvalidArgumentError("pauli_sums must be rank 2.")
        sums = [[parse_proto(s, _pb.PauliSum) for s in row]
                for row in pauli_sums]
        if len(sums) != len(progs):
            raise InvalidArgumentError(
                "Number of circuits and PauliSums do not match.")
    nq = [resolve_qubit_ids(p, None if sums is None else sums[i])
          for i, p in enumerate(progs)]
    maps = _symbol_maps(symbol_names, symbol_values)
    if len(maps) != len(progs):
        raise InvalidArgumentError(
            "Number of circuits and symbol_values do not match.")
    circuits = [circuit_from_program(p, m, n)
                for p, m, n in zip(progs, maps, nq)]
    return progs, sums, nq, maps, circuits


def simulate_expectation(programs, symbol_names, symbol_values, pauli_sums,
                         backend="c", threads=1):
    """TfqSimulateExpectation (tfq_simulate_expectation_op.cc:50-250)."""
    progs, sums, nq, maps, circuits = _prologue(
        programs, symbol_names, symbol_values, pauli_sums)
    B = len(progs)
    M = len(sums[0]) if B else 0
    out = np.zeros((B, M), dtype=np.float32)
    jobs, job_rows = [], []
    for i in range(B):
        if len(circuits[i]) == 0:
            out[i, :] = -2.0
            continue
        steps = forward_steps(circuits[i])
        idents = []
        for j, ps in enumerate(sums[i]):
            st, idn = _strip_ident(expectation_steps(ps, j))
            steps += st
            idents += idn
        jobs.append((steps, nq[i], M, idents))
        job_rows.append(i)
    results = _run_jobs(jobs, backend, threads)
    for i, (res, job) in zip(job_rows, zip(results, jobs)):
        vals = res[:M].copy()
        for _, slot, c in job[3]:
            vals[slot] += float(c)
        out[i, :] = vals.astype(np.float32)
    return out


def _run_jobs(jobs, backend, threads):
    if not jobs:
        return []
    if backend in ("numpy", "numpy128"):
        return [_run(s, n, k, backend) for s, n, k, *_ in jobs]
    return run_batch_c([j[0] for j in jobs], [j[1] for j in jobs],
                       [j[2] for j in jobs], threads)


def simulate_state(programs, symbol_names, symbol_values, backend="c"):
    """TfqSimulateState (tfq_simulate_state_op.cc:48-216)."""
    progs, _, nq, maps, circuits = _prologue(programs, symbol_names,
                                             symbol_values)
    B = len(progs)
    nmax = max(nq) if B else 0
    out = np.full((B, 2 ** nmax), np.complex64(-2), dtype=np.complex64)
    for i in range(B):
        _, st = _run(forward_steps(circuits[i]), nq[i], 0, backend, True)
        out[i, :2 ** nq[i]] = st[:2 ** nq[i]]
    return out


def _final_state(circuit, n, backend):
    _, st = _run(forward_steps(circuit), n, 0, backend, True)
    return st


def simulate_samples(programs, symbol_names, symbol_values, num_samples,
                     uniforms=None, seed=0, backend="c"):
    """TfqSimulateSamples (tfq_simulate_samples_op.cc:53-252). `uniforms`
    [B,S] in [0,1) (sorted ascending per row before use, as qsim sorts its
    draws); else Philox(seed) stream (shot,row,0,0)."""
    progs, _, nq, maps, circuits = _prologue(programs, symbol_names,
                                             symbol_values)
    ns = np.asarray(num_samples).reshape(-1)
    if len(ns) != 1:
        raise InvalidArgumentError("num_samples must contain 1 element.")
    S = int(ns[0])
    B = len(progs)
    nmax = max(nq) if B else 0
    out = np.zeros((B, S, nmax), dtype=np.int8)
    if S == 0:
        return out
    for i in range(B):
        st = _final_state(circuits[i], nq[i], backend)
        u = (np.asarray(uniforms[i], dtype=np.float64) if uniforms is not None
             else philox_uniforms(seed, i, 0, 0, S))
        idx = sample_tree(st[:2 ** max(nq[i], 1)], np.sort(u))
        for q in range(nq[i]):
            out[i, :, nmax - 1 - q] = (idx >> q) & 1
        out[i, :, :nmax - nq[i]] = -2
    return out


def samples_from_states(states, nq, num_samples, uniforms):
    """The sampling half of TfqSimulateSamples applied to GIVEN final states
    (rows of a TfqSimulateState output, -2 padded): isolates the sampler so
    that 'bit-exact for identical uniforms' can be checked without float32
    state round-off moving a CDF boundary across a uniform."""
    B = len(nq)
    nmax = max(nq) if B else 0
    S = int(num_samples)
    out = np.zeros((B, S, nmax), dtype=np.int8)
    for i in range(B):
        if nq[i] == 0:
            out[i] = -2
            continue
        st = np.asarray(states[i][:2 ** nq[i]], dtype=np.complex64)
        idx = sample_tree(st, np.sort(np.asarray(uniforms[i], dtype=np.float64)))
        for q in range(nq[i]):
            out[i, :, nmax - 1 - q] = (idx >> q) & 1
        out[i, :, :nmax - nq[i]] = -2
    return out


def num_qubits(programs):
    """Resolved qubit count per program (program_resolution.cc:90-186)."""
    return [resolve_qubit_ids(parse_proto(p, _pb.Program)) for p in programs]


def simulate_sampled_expectation(programs, symbol_names, symbol_values,
                                 pauli_sums, num_samples, uniforms=None,
                                 seed=0, backend="c"):
    """TfqSimulateSampledExpectation (..sampled_expectation_op.cc:54-306,
    util_qsim.h:199-270). Uniform stream for (row i, op j, term t) =
    Philox(seed) counter (shot, i, j, t), or `uniforms[i][j][t]` if given."""
    progs, sums, nq, maps, circuits = _prologue(
        programs, symbol_names, symbol_values, pauli_sums)
    nsamp = np.asarray(num_samples)
    if nsamp.ndim != 2:
        raise InvalidArgumentError("num_samples must be rank 2. Got rank %d."
                                   % nsamp.ndim)
    if (nsamp < 1).any():
        raise InvalidArgumentError(
            "Each element of num_samples must be greater than 0.")
    B = len(progs)
    M = len(sums[0]) if B else 0
    out = np.zeros((B, M), dtype=np.float32)
    for i in range(B):
        if len(circuits[i]) == 0:
            out[i, :] = -2.0
            continue
        n = nq[i]
        psi = _final_state(circuits[i], n, backend)
        for j, ps in enumerate(sums[i]):
            S = int(nsamp[i, j])
            e = F32(0)
            for t, term in enumerate(ps.terms):
                if len(term.paulis) == 0:
                    e = F32(e + F32(term.coefficient_real))
                    continue
                phi = psi.copy()
                for g in _zbasis_gates(term):
                    _np_apply(phi, max(n, 1), g.qubits, g.matrix)
                u = (np.asarray(uniforms[i][j][t], dtype=np.float64)[:S]
                     if uniforms is not None
                     else philox_uniforms(seed, i, j, t, S))
                idx = sample_tree(phi, u)
                mask = 0
                for p in term.paulis:
                    mask |= 1 << (n - int(p.qubit_id) - 1)
                par = np.zeros(S, dtype=np.int64)
                m = idx & mask
                while m.any():
                    par ^= m & 1
                    m >>= 1
                total = int(np.sum(1 - 2 * par))
                e = F32(e + F32(F32(F32(total) * F32(term.coefficient_real))
                                / F32(S)))
            out[i, j] = e
    return out


def adjoint_gradient(programs, symbol_names, symbol_values, pauli_sums,
                     downstream_grads, backend="c", threads=1, out_dtype=np.float32):
    """TfqAdjointGradient (tfq_adj_grad_op.cc:51-390).  `out_dtype=np.float64`
    with backend "numpy128" gives the double-state yardstick un-rounded."""
    progs, sums, nq, maps, circuits = _prologue(
        programs, symbol_names, symbol_values, pauli_sums)
    down = np.asarray(downstream_grads, dtype=np.float32)
    B = len(progs)
    if down.ndim != 2:
        raise InvalidArgumentError("downstream_grads must be rank 2.")
    if down.shape[0] != B:
        raise InvalidArgumentError(
            "Number of gradients and circuits do not match.")
    M = len(sums[0]) if B else 0
    if down.shape[1] != M:
        raise InvalidArgumentError(
            "Number of gradients and pauli sum dimension do not match.")
    names = [s.decode() if isinstance(s, bytes) else s for s in symbol_names]
    P = len(names)
    sym_col = {nm: maps[0][nm][0] for nm in names} if B else {}
    out = np.zeros((B, P), dtype=out_dtype)
    jobs, rows = [], []
    for i in range(B):
        if len(circuits[i]) == 0:
            continue
        gg = gradient_gates(circuits[i])
        steps = adjoint_steps(circuits[i], gg, sums[i], down[i], sym_col)
        jobs.append((steps, nq[i], P))
        rows.append(i)
    for i, res in zip(rows, _run_jobs(jobs, backend, threads)):
        out[i, :] = res[:P].astype(out_dtype)
    return out


def inner_product(programs, symbol_names, symbol_values, other_programs,
                  backend="c"):
    """TfqInnerProduct (math_ops/tfq_inner_product.cc:45-292):
    out[i][j] = <psi_i | phi_ij>, fp64 accumulation, (1, 0) for an empty
    programs[i]; paired circuits are symbol free and must use exactly the
    reference circuit's qubits (program_resolution.cc:188-311)."""
    if np.ndim(programs) != 1:
        raise InvalidArgumentError("programs must be rank 1. Got rank %d."
                                   % np.ndim(programs))
    progs = [parse_proto(p, _pb.Program) for p in programs]
    B = len(progs)
    if len(other_programs) != B:
        raise InvalidArgumentError(
            "programs and other_programs batch dimension do not match. Foud: "
            "%d and %d" % (B, len(other_programs)))
    K = len(other_programs[0]) if B else 0
    maps = _symbol_maps(symbol_names, symbol_values)
    if len(maps) != B:
        raise InvalidArgumentError(
            "Number of circuits and symbol_values do not match.")
    out = np.zeros((B, K), dtype=np.complex64)
    for i in range(B):
        others = [parse_proto(o, _pb.Program) for o in other_programs[i]]
        ids = {}
        if len(progs[i].circuit.moments):
            seen = set()
            for m in progs[i].circuit.moments:
                for op in m.operations:
                    for q in op.qubits:
                        _register_qubits(q.id, seen)
                    _register_qubits(
                        op.args["control_qubits"].arg_value.string_value, seen)
            ids = {k[1]: str(j) for j, k in enumerate(sorted(seen))}
        n = resolve_qubit_ids(progs[i])
        if n == 0:
            out[i, :] = 1.0
            continue
        psi = _final_state(circuit_from_program(progs[i], maps[i], n), n, backend)
        for j, o in enumerate(others):
            unvisited = set(ids)
            for m in o.circuit.moments:
                for op in m.operations:
                    for arg in op.args.values():
                        if arg.symbol:
                            raise InvalidArgumentError(
                                "Found symbols in other_programs.No symbols "
                                "are allowed in these circuits.")
                    for q in op.qubits:
                        unvisited.discard(q.id)
                        if q.id not in ids:
                            raise InvalidArgumentError(
                                "A paired circuit contains qubits not found "
                                "in reference circuit.")
                        q.id = ids[q.id]
                    cq = op.args["control_qubits"].arg_value.string_value
                    if cq:
                        toks = cq.split(",")
                        for t in toks:
                            unvisited.discard(t)
                            if t not in ids:
                                raise InvalidArgumentError(
                                    "A paired circuit contains qubits not "
                                    "found in reference circuit.")
                        op.args["control_qubits"].arg_value.string_value = \
                            ",".join(ids[t] for t in toks)
            if unvisited:
                raise InvalidArgumentError(
                    "A reference circuit contains qubits not found in paired "
                    "circuit.")
            phi = _final_state(circuit_from_program(o, {}, n), n, backend)
            out[i, j] = np.complex64(np.vdot(psi[:2 ** n].astype(np.complex128),
                                             phi[:2 ** n].astype(np.complex128)))
    return out


def _reference_qubit_ids(prog):
    """Sorted qubit ids of a reference program -> decimal index strings
    (program_resolution.cc:188-311), or {} for an empty program."""
    if not len(prog.circuit.moments):
        return {}
    seen = set()
    for m in prog.circuit.moments:
        for op in m.operations:
            for q in op.qubits:
                _register_qubits(q.id, seen)
            _register_qubits(
                op.args["control_qubits"].arg_value.string_value, seen)
    return {k[1]: str(j) for j, k in enumerate(sorted(seen))}


def _resolve_paired(o, ids):
    """Rewrite a paired (symbol free) program's qubit ids against the
    reference circuit's map, with the reference's error strings."""
    unvisited = set(ids)
    for m in o.circuit.moments:
        for op in m.operations:
            for arg in op.args.values():
                if arg.symbol:
                    raise InvalidArgumentError(
                        "Found symbols in other_programs.No symbols "
                        "are allowed in these circuits.")
            for q in op.qubits:
                unvisited.discard(q.id)
                if q.id not in ids:
                    raise InvalidArgumentError(
                        "A paired circuit contains qubits not found "
                        "in reference circuit.")
                q.id = ids[q.id]
            cq = op.args["control_qubits"].arg_value.string_value
            if cq:
                toks = cq.split(",")
                for t in toks:
                    unvisited.discard(t)
                    if t not in ids:
                        raise InvalidArgumentError(
                            "A paired circuit contains qubits not "
                            "found in reference circuit.")
                op.args["control_qubits"].arg_value.string_value = \
                    ",".join(ids[t] for t in toks)
    if unvisited:
        raise InvalidArgumentError(
            "A reference circuit contains qubits not found in paired "
            "circuit.")


def inner_product_grad(programs, symbol_names, symbol_values, other_programs,
                       downstream_grads, backend="c"):
    """TfqInnerProductGrad (math_ops/tfq_inner_product_grad.cc:46-501), what
    the op returns (the Python wrapper conjugates it):
      lam = sum_j downstream[i][j] |phi_ij>   (util_qsim.h:422-442)
      reverse sweep as in the adjoint op, but with the complex
      out[i][col] += <dG psi' | lam>  (fp64 inner product, complex64 sum)
    at every gradient gate (:255-310)."""
    if np.ndim(programs) != 1:
        raise InvalidArgumentError("programs must be rank 1. Got rank %d."
                                   % np.ndim(programs))
    names = [s.decode() if isinstance(s, bytes) else s for s in symbol_names]
    P = len(names)
    if P == 0:
        raise InvalidArgumentError(
            "The number of symbols must be a positive integer, got 0 symbols.")
    progs = [parse_proto(p, _pb.Program) for p in programs]
    B = len(progs)
    if len(other_programs) != B:
        raise InvalidArgumentError(
            "programs and other_programs batch dimension do not match. Foud: "
            "%d and %d" % (B, len(other_programs)))
    K = len(other_programs[0]) if B else 0
    down = np.asarray(downstream_grads, dtype=np.float32)
    if down.ndim != 2 or down.shape[0] != B:
        raise InvalidArgumentError(
            "Number of gradients and circuits do not match.")
    if down.shape[1] != K:
        raise InvalidArgumentError(
            "Number of gradients and other_programs do not match.")
    maps = _symbol_maps(symbol_names, symbol_values)
    if len(maps) != B:
        raise InvalidArgumentError(
            "Number of circuits and symbol_values do not match.")
    out = np.zeros((B, P), dtype=np.complex64)
    for i in range(B):
        ids = _reference_qubit_ids(progs[i])
        others = [parse_proto(o, _pb.Program) for o in other_programs[i]]
        n = resolve_qubit_ids(progs[i])
        if n == 0:
            continue
        gates = circuit_from_program(progs[i], maps[i], n)
        sym_col = {nm: maps[i][nm][0] for nm in names}
        sv = _final_state(gates, n, backend)[:2 ** n].astype(np.complex64).copy()
        lam = np.zeros(2 ** n, dtype=np.complex64)
        for j, o in enumerate(others):
            _resolve_paired(o, ids)
            phi = _final_state(circuit_from_program(o, {}, n), n,
                               backend)[:2 ** n].astype(np.complex64)
            c = F32(down[i, j])
            scaled = (phi.real * c + 1j * (phi.imag * c)).astype(np.complex64)
            lam = _c64_f32(np.add, lam, scaled)
        grads = gradient_gates(gates)
        bounds = [gg.index for gg in grads]
        segs, left = [], 0
        for b in bounds:
            segs.append(basic_fuse(gates[left:b]))
            left = b + 1
        segs.append(basic_fuse(gates[left:]))
        for j in range(len(segs) - 1, -1, -1):
            for f in reversed(segs[j]):
                md = dagger(f.matrix)
                _np_apply(sv, n, f.qubits, md, f.controls, f.cvalues)
                _np_apply(lam, n, f.qubits, md, f.controls, f.cvalues)
            if j == 0:
                break
            gg = grads[j - 1]
            cur = gates[gg.index]
            cd = dagger(cur.matrix)
            _np_apply(sv, n, cur.qubits, cd, cur.controls, cur.cvalues)
            for sym, dm in zip(gg.symbols, gg.matrices):
                s2 = sv.copy()
                if cur.controls:
                    psi = s2.reshape((2,) * n)
                    keep = np.zeros((2,) * n, dtype=bool)
                    idx = [slice(None)] * n
                    for a, v in zip(cur.controls, cur.cvalues):
                        idx[a] = v
                    keep[tuple(idx)] = True
                    psi[~keep] = 0
                _np_apply(s2, n, cur.qubits, dm)
                r = np.vdot(s2.astype(np.complex128), lam.astype(np.complex128))
                col = sym_col[sym]
                out[i, col] = np.complex64(
                    out[i, col] + np.complex64(complex(F32(r.real), F32(r.imag))))
            _np_apply(lam, n, cur.qubits, cd, cur.controls, cur.cvalues)
    return out


# ==========================================================================
# (6) next-row N2: noisy trajectory ops
#     TfqNoisyExpectation        core/ops/noise/tfq_noisy_expectation.cc:57-391
#     TfqNoisySampledExpectation core/ops/noise/tfq_noisy_sampled_expectation.cc:57-404
#     TfqNoisySamples            core/ops/noise/tfq_noisy_samples.cc:54-321
#     NoisyQsimCircuitFromProgram + channel builders
#                                core/src/circuit_parser_qsim.cc:598-826
# The channel definitions and the trajectory step live in qsim v0.21.0
# (lib/channels_cirq.h, lib/qtrajectory.h), which is NOT under /root/reference:
# they are restated here from qsim's published code.  PARITY UNPINNED for this
# section: the reference seeds qsim's std::mt19937 from a non-deterministic
# Philox stream and its tests only compare with cirq statistically
# (noise/tfq_noisy_expectation_test.py), so there is no golden trajectory; the
# "same uniforms -> same trajectory" contract below (one uniform per noise
# channel, in program order) is this repository's.
# ==========================================================================

def _kraus(unitary, prob, m):
    return (bool(unitary), float(prob), np.asarray(m, dtype=np.complex64).reshape(2, 2))


def channel_kraus(gid: str, a: dict):
    """[(unitary, probability or lower bound, 2x2 matrix)] in qsim's order
    (lib/channels_cirq.h).  For a unitary Kraus operator `prob` is its
    probability and the matrix is the UNSCALED unitary; for a non-unitary one
    it is the lower bound min eig(K^dagger K) that qsim accepts without
    computing a norm, and the matrix is K itself."""
    I = [[1, 0], [0, 1]]
    X = [[0, 1], [1, 0]]
    Y = [[0, -1j], [1j, 0]]
    Z = [[1, 0], [0, -1]]
    if gid == "ADP":
        px, py, pz = float(a["p_x"]), float(a["p_y"]), float(a["p_z"])
        return [_kraus(1, 1 - px - py - pz, I), _kraus(1, px, X), _kraus(1, py, Y),
                _kraus(1, pz, Z)]
    if gid == "DP":
        p = float(a["p"])
        return [_kraus(1, 1 - p, I), _kraus(1, p / 3, X), _kraus(1, p / 3, Y),
                _kraus(1, p / 3, Z)]
    if gid == "BF":
        p = float(a["p"])
        return [_kraus(1, 1 - p, I), _kraus(1, p, X)]
    if gid == "PF":
        p = float(a["p"])
        return [_kraus(1, 1 - p, I), _kraus(1, p, Z)]
    if gid == "AD":
        g = float(a["gamma"])
        return [_kraus(0, 1 - g, [[1, 0], [0, np.sqrt(1 - g)]]),
                _kraus(0, 0.0, [[0, np.sqrt(g)], [0, 0]])]
    if gid == "PD":
        g = float(a["gamma"])
        return [_kraus(0, 1 - g, [[1, 0], [0, np.sqrt(1 - g)]]),
                _kraus(0, 0.0, [[0, 0], [0, np.sqrt(g)]])]
    if gid == "RST":
        return [_kraus(0, 0.0, [[1, 0], [0, 0]]), _kraus(0, 0.0, [[0, 1], [0, 0]])]
    if gid == "GAD":
        p, g = float(a["p"]), float(a["gamma"])
        return [_kraus(0, p * (1 - g), [[np.sqrt(p), 0], [0, np.sqrt(p * (1 - g))]]),
                _kraus(0, (1 - p) * (1 - g), [[np.sqrt((1 - p) * (1 - g)), 0],
                                              [0, np.sqrt(1 - p)]]),
                _kraus(0, 0.0, [[0, np.sqrt(p * g)], [0, 0]]),
                _kraus(0, 0.0, [[0, 0], [np.sqrt((1 - p) * g), 0]])]
    raise InvalidArgumentError("Could not parse channel id: " + gid)


_CHANNEL_ARGS = {"DP": ("p",), "ADP": ("p_x", "p_y", "p_z"), "GAD": ("p", "gamma"),
                 "AD": ("gamma",), "RST": (), "PD": ("gamma",), "PF": ("p",),
                 "BF": ("p",)}


def noisy_circuit_from_program(program, smap, n):
    """NoisyQsimCircuitFromProgram (circuit_parser_qsim.cc:773-826): a list of
    ("gate", Gate) and ("channel", axis, kraus) in moment order."""
    if n <= 0:
        return []
    items = []
    for moment in program.circuit.moments:
        for op in moment.operations:
            gid = op.gate.id
            if gid in _ALL_IDS:
                items.append(("gate", build_gate(op, smap)))
                continue
            if gid not in _CHANNEL_ARGS:
                raise InvalidArgumentError("Could not parse channel id: " + gid)
            args = {}
            for name in _CHANNEL_ARGS[gid]:
                args[name] = _arg(op, name, {})
            items.append(("channel", int(op.qubits[0].id), channel_kraus(gid, args)))
    return items


def count_channels(items):
    return sum(1 for it in items if it[0] == "channel")


def run_trajectory(items, n, uniforms):
    """QuantumTrajectorySimulator::RunOnce (qsim lib/qtrajectory.h) with one
    uniform per noise channel: first the cumulative (lower-bound)
    probabilities; if r is beyond them, the true probabilities
    <psi|K^dagger K|psi> of the non-unitary operators on the normalised state,
    in order, with the most probable one as the round-off fallback.  Returns
    the normalised complex64 state."""
    psi = np.zeros(2 ** max(n, 1), dtype=np.complex64)
    psi[0] = 1
    c = 0
    for it in items:
        if it[0] == "gate":
            g = it[1]
            _np_apply(psi, max(n, 1), g.qubits, g.matrix, g.controls, g.cvalues)
            continue
        _, axis, kraus = it
        r = float(np.float32(uniforms[c]))
        c += 1
        cp, chosen = 0.0, None
        for k, (unitary, prob, m) in enumerate(kraus):
            cp += prob
            if r < cp:
                chosen = k
                break
        if chosen is None:
            nrm = np.sqrt(np.vdot(psi, psi).real)
            psi = (psi / np.float32(nrm)).astype(np.complex64)
            view = np.moveaxis(psi.reshape((2,) * max(n, 1)), axis, 0).reshape(2, -1)
            pop = (np.abs(view.astype(np.complex128)) ** 2).sum(axis=1)
            best, best_p = 0, -1.0
            for k, (unitary, prob, m) in enumerate(kraus):
                if unitary:
                    continue
                kd = (m.conj().T @ m).astype(np.complex128)
                pk = float(kd[0, 0].real * pop[0] + kd[1, 1].real * pop[1])
                if pk > best_p:
                    best, best_p = k, pk
                cp += pk - prob
                if r < cp or k == len(kraus) - 1:
                    chosen = k if r < cp else best
                    break
        m = kraus[chosen][2]
        _np_apply(psi, max(n, 1), (axis,), m)
        if not kraus[chosen][0]:
            nrm = np.sqrt(np.vdot(psi, psi).real)
            psi = (psi / np.float32(nrm)).astype(np.complex64)
    return psi


NOISE_STREAM = 0x6E6F6973          # "nois": stream_b of the channel uniforms


def channel_uniforms(seed, row, trajectory, count):
    """Uniform c of (row, trajectory): Philox counter (c, row, trajectory,
    NOISE_STREAM), rounded to float32 (it travels in the float parameter row
    of the device)."""
    return philox_uniforms(seed, row, trajectory, NOISE_STREAM, count).astype(np.float32)


def _expectation_of_state(psi, n, ps):
    """ComputeExpectationQsim (util_qsim.h:142-188) on a given state."""
    e = F32(0)
    for term in ps.terms:
        if len(term.paulis) == 0:
            e = F32(e + F32(term.coefficient_real))
            continue
        phi = psi.copy()
        for g in _pauli_term_gates(term):
            _np_apply(phi, max(n, 1), g.qubits, g.matrix)
        v = np.vdot(psi.astype(np.complex128), phi.astype(np.complex128)).real
        e = F32(e + F32(F32(term.coefficient_real) * F32(v)))
    return e


def noisy_expectation(programs, symbol_names, symbol_values, pauli_sums, num_samples,
                      uniforms=None, seed=0):
    """TfqNoisyExpectation (tfq_noisy_expectation.cc:57-391): out[i, j] = mean
    over the first num_samples[i][j] trajectories of row i of the exact
    <psi_t| O_j |psi_t>.  `uniforms[i][t][c]` overrides the Philox stream."""
    progs, sums, nq, maps, _ = _noisy_prologue(programs, symbol_names, symbol_values,
                                               pauli_sums)
    ns = _check_num_samples(num_samples, sums)
    B = len(progs)
    M = len(sums[0]) if B else 0
    out = np.zeros((B, M), dtype=np.float32)
    for i in range(B):
        items = noisy_circuit_from_program(progs[i], maps[i], nq[i])
        if not items:
            out[i, :] = -2.0
            continue
        C = count_channels(items)
        T = int(ns[i].max())
        acc = np.zeros(M, dtype=np.float64)
        for t in range(T):
            u = (np.asarray(uniforms[i][t], dtype=np.float32) if uniforms is not None
                 else channel_uniforms(seed, i, t, C))
            psi = run_trajectory(items, nq[i], u)
            for j in range(M):
                if t < ns[i, j]:
                    acc[j] += float(_expectation_of_state(psi, nq[i], sums[i][j]))
        out[i, :] = (acc / ns[i]).astype(np.float32)
    return out


def _noisy_prologue(programs, symbol_names, symbol_values, pauli_sums):
    if np.ndim(programs) != 1:
        raise InvalidArgumentError("programs must be rank 1. Got rank %d."
                                   % np.ndim(programs))
    progs = [parse_proto(p, _pb.Program) for p in programs]
    sums = None
    if pauli_sums is not None:
        sums = [[parse_proto(s, _pb.PauliSum) for s in row] for row in pauli_sums]
        if len(sums) != len(progs):
            raise InvalidArgumentError("Number of circuits and PauliSums do not match.")
    nq = [resolve_qubit_ids(p, None if sums is None else sums[i])
          for i, p in enumerate(progs)]
    maps = _symbol_maps(symbol_names, symbol_values)
    if len(maps) != len(progs):
        raise InvalidArgumentError("Number of circuits and symbol_values do not match.")
    return progs, sums, nq, maps, None


def _check_num_samples(num_samples, sums):
    ns = np.asarray(num_samples)
    if ns.ndim != 2:
        raise InvalidArgumentError("num_samples must be rank 2. Got rank %d." % ns.ndim)
    if ns.shape[0] != len(sums):
        raise InvalidArgumentError(
            "Dimension 0 of num_samples and pauli_sums do not match.")
    if len(sums) and ns.shape[1] != len(sums[0]):
        raise InvalidArgumentError(
            "Dimension 1 of num_samples and pauli_sums do not match.")
    if (ns < 1).any():
        raise InvalidArgumentError("Each element of num_samples must be greater than 0.")
    return ns.astype(np.int64)


SAMPLE_STREAM = 0x73616D70         # "samp": measurement uniforms of noisy ops


def noisy_samples(programs, symbol_names, symbol_values, num_samples, uniforms=None,
                  measure_uniforms=None, seed=0):
    """TfqNoisySamples (tfq_noisy_samples.cc:54-321): shot s of row i is ONE
    bitstring measured at the end of trajectory s (a terminal measurement of
    every qubit).  int8 [B, S, nmax], -2 padded on the left.  Measurement
    uniform of (i, s): Philox counter (0, i, s, SAMPLE_STREAM)."""
    progs, _, nq, maps, _ = _noisy_prologue(programs, symbol_names, symbol_values, None)
    S = int(np.asarray(num_samples).reshape(-1)[0])
    B = len(progs)
    nmax = max(nq) if B else 0
    out = np.zeros((B, S, nmax), dtype=np.int8)
    for i in range(B):
        items = noisy_circuit_from_program(progs[i], maps[i], nq[i])
        C = count_channels(items)
        for s in range(S):
            if nq[i] == 0:
                out[i, s] = -2
                continue
            u = (np.asarray(uniforms[i][s], dtype=np.float32) if uniforms is not None
                 else channel_uniforms(seed, i, s, C))
            psi = run_trajectory(items, nq[i], u)
            um = (np.asarray([measure_uniforms[i][s]], dtype=np.float64)
                  if measure_uniforms is not None
                  else philox_uniforms(seed, i, s, SAMPLE_STREAM, 1))
            idx = int(sample_tree(psi, um)[0])
            for q in range(nq[i]):
                out[i, s, nmax - 1 - q] = (idx >> q) & 1
            out[i, s, :nmax - nq[i]] = -2
    return out


def noisy_sampled_expectation(programs, symbol_names, symbol_values, pauli_sums,
                              num_samples, uniforms=None, seed=0):
    """TfqNoisySampledExpectation (tfq_noisy_sampled_expectation.cc:57-404):
    like noisy_expectation, but every trajectory contributes ONE shot per term
    (ComputeSampledExpectationQsim with num_samples = 1, :238-240).  Shot
    uniform of (row i, trajectory t, op j, term k): Philox counter
    (k, i, t, SAMPLE_STREAM + 1 + j)."""
    progs, sums, nq, maps, _ = _noisy_prologue(programs, symbol_names, symbol_values,
                                               pauli_sums)
    ns = _check_num_samples(num_samples, sums)
    B = len(progs)
    M = len(sums[0]) if B else 0
    out = np.zeros((B, M), dtype=np.float32)
    for i in range(B):
        items = noisy_circuit_from_program(progs[i], maps[i], nq[i])
        if not items:
            out[i, :] = -2.0
            continue
        C = count_channels(items)
        n = nq[i]
        T = int(ns[i].max())
        acc = np.zeros(M, dtype=np.float64)
        for t in range(T):
            u = (np.asarray(uniforms[i][t], dtype=np.float32) if uniforms is not None
                 else channel_uniforms(seed, i, t, C))
            psi = run_trajectory(items, n, u)
            for j in range(M):
                if t >= ns[i, j]:
                    continue
                terms = sums[i][j].terms
                us = philox_uniforms(seed, i, t, SAMPLE_STREAM + 1 + j, max(len(terms), 1))
                e = F32(0)
                for k, term in enumerate(terms):
                    if len(term.paulis) == 0:
                        e = F32(e + F32(term.coefficient_real))
                        continue
                    phi = psi.copy()
                    for g in _zbasis_gates(term):
                        _np_apply(phi, max(n, 1), g.qubits, g.matrix)
                    idx = int(sample_tree(phi, us[k:k + 1])[0])
                    mask = 0
                    for p in term.paulis:
                        mask |= 1 << (n - int(p.qubit_id) - 1)
                    par = bin(idx & mask).count("1") & 1
                    e = F32(e + F32(F32(1 - 2 * par) * F32(term.coefficient_real)))
                acc[j] += float(e)
        out[i, :] = (acc / ns[i]).astype(np.float32)
    return out


# ==========================================================================
# (7) next-row N4: TfqCalculateUnitary (tfq_calculate_unitary_op.cc:47-164)
# ==========================================================================

def calculate_unitary(programs, symbol_names, symbol_values):
    """out[i, j, k] = <j| U_i |k>; (-2, 0) outside a smaller circuit's block.
    The columns are the circuit applied gate by gate (fused like the op's
    fused_circuits) to the basis states."""
    progs, _, nq, maps, circuits = _prologue(programs, symbol_names, symbol_values)
    B = len(progs)
    nmax = max(nq) if B else 0
    D = 2 ** nmax
    out = np.full((B, D, D), np.complex64(-2), dtype=np.complex64)
    for i in range(B):
        n = nq[i]
        d = 2 ** n
        fused = basic_fuse(circuits[i])
        for k in range(d):
            v = np.zeros(2 ** max(n, 1), dtype=np.complex64)
            v[k] = 1
            for g in fused:
                _np_apply(v, max(n, 1), g.qubits, g.matrix, g.controls, g.cvalues)
            out[i, :d, k] = v[:d]
    return out
