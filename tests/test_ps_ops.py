"""Parameter-shift helper ops: TfqPsDecompose,
TfqPsSymbolReplace, TfqPsWeightsFromSymbols.  Host-only ops: every test runs
on CPU.  The cases restate the reference's
core/ops/tfq_ps_util_ops_test.py (decompositions checked through the
unitary, :71-124, :266-363; symbol replacement :368-669; weights :674-853)
with this repo's generators and the oracle's unitary in place of cirq's."""
import numpy as np
import pytest

from oracle import tfq_oracle as orc
from quantum_b200 import circuits as cq
from quantum_b200 import ops, proto


def _parse(b):
    p = proto.Program()
    p.ParseFromString(bytes(b))
    return p


def _ops_of(b):
    """[(moment, gate id, {arg: value | 'sym:<name>'}, [qubit ids])]"""
    out = []
    for j, m in enumerate(_parse(b).circuit.moments):
        for op in m.operations:
            args = {}
            for k, a in op.args.items():
                kind = a.WhichOneof("arg")
                if kind == "symbol":
                    args[k] = "sym:" + a.symbol
                else:
                    which = a.arg_value.WhichOneof("arg_value")
                    args[k] = getattr(a.arg_value, which)
            out.append((j, op.gate.id, args, [q.id for q in op.qubits]))
    return out


def _unitary(prog, names, vals):
    u = orc.calculate_unitary([prog], names, np.asarray([vals], np.float32))
    return u[0]


# ---------------------------------------------------------------- decompose
@pytest.mark.parametrize("case", ["iswap", "phased_x", "fsim", "phased_iswap"])
def test_decompose_single_gate_keeps_the_unitary(case):
    """tfq_ps_util_ops_test.py:71-124: the decomposed program has the
    unitary of the original one for random symbol values."""
    rng = np.random.default_rng({"iswap": 1, "phased_x": 2, "fsim": 3, "phased_iswap": 4}[case])
    q = [cq.grid(0, i) for i in range(2)]
    c1, c2 = float(rng.random()), float(rng.random())
    if case == "iswap":
        m = [[cq.ISWAP(q[0], q[1], "t", scalar=c1)]]
    elif case == "phased_x":
        m = [[cq.PhasedX(q[0], "r", "t", phase_scalar=c1, scalar=c2)]]
    elif case == "fsim":
        m = [[cq.FSim(q[0], q[1], "t", "r", theta_scalar=c1, phi_scalar=c2)]]
    else:
        m = [[cq.PhasedISwap(q[0], q[1], "r", "t", phase_scalar=c1, scalar=c2)]]
    prog = cq.serialize(m)
    out = ops.tfq_ps_decompose([prog])
    assert out.shape == (1,)
    ids = [o[1] for o in _ops_of(out[0])]
    assert not set(ids) & {"ISP", "PXP", "FSIM", "PISP"}
    vals = rng.random(2)
    a = _unitary(prog, ["t", "r"], vals)
    b = _unitary(out[0], ["t", "r"], vals)
    np.testing.assert_allclose(a, b, atol=1e-5)


def test_decompose_leaves_literal_gates_and_keeps_moments():
    """:266-363: only parameterised composite gates are decomposed; their
    extra moments follow the moment they came from; language and scheduling
    strategy are what TFQ's serializer writes."""
    q = [cq.grid(0, i) for i in range(6)]
    m = [
        [cq.H(x) for x in q],
        [cq.X(q[4]), cq.PhasedX(q[5], "t", 0.4), cq.ISWAP(q[0], q[1], "t", scalar=0.7),
         cq.FSim(q[2], q[3], "t", "r", theta_scalar=0.3, phi_scalar=0.9)],
        [cq.ISWAP(q[0], q[1], 0.25), cq.FSim(q[2], q[3], 0.1, 0.2), cq.PhasedX(q[4], 0.3, 0.5)],
        [cq.H(x) for x in q],
    ]
    prog = cq.serialize(m)
    out = ops.tfq_ps_decompose([prog, prog])
    p = _parse(out[1])
    assert p.language.gate_set == "tfq_gate_set"
    assert p.circuit.scheduling_strategy == 1            # MOMENT_BY_MOMENT
    got = _ops_of(out[0])
    # moment 1 keeps 4 operations and is followed by 2 extra moments
    # (PXP / FSIM need two, ISP one); moment 2 is untouched
    assert len(p.circuit.moments) == 6
    by_moment = [[o[1] for o in got if o[0] == j] for j in range(6)]
    assert by_moment[0] == ["HP"] * 6 and by_moment[5] == ["HP"] * 6
    assert sorted(by_moment[1]) == sorted(["XP", "ZP", "XXP", "XXP"])
    assert sorted(by_moment[2]) == sorted(["XP", "YYP", "YYP"])
    assert sorted(by_moment[3]) == sorted(["ZP", "CZP"])
    assert by_moment[4] == ["ISP", "FSIM", "PXP"]
    vals = np.array([0.37, 0.81])
    np.testing.assert_allclose(_unitary(prog, ["t", "r"], vals),
                               _unitary(out[0], ["t", "r"], vals), atol=2e-5)


def test_decompose_controls_and_text_format():
    """Control metadata is copied onto every factor; text-format programs are
    accepted like everywhere else (parse_context.cc:41-56)."""
    q = [cq.grid(0, i) for i in range(3)]
    m = [[cq.ISWAP(q[0], q[1], "t", scalar=0.6).controlled_by([q[2]], [0])]]
    out = ops.tfq_ps_decompose([cq.serialize(m), cq.serialize_text(m)])
    assert bytes(out[0]) == bytes(out[1])
    for o in _ops_of(out[0]):
        assert o[2]["control_values"] == "0" and o[2]["control_qubits"] != ""
    vals = np.array([0.23])
    np.testing.assert_allclose(_unitary(cq.serialize(m), ["t"], vals),
                               _unitary(out[0], ["t"], vals), atol=1e-5)


# ----------------------------------------------------------- symbol replace
def test_symbol_replace_simple_and_padding():
    """:368-398, :480-558: one copy per occurrence, that occurrence renamed,
    everything else untouched; shorter lists padded with empty programs."""
    q = [cq.grid(0, i) for i in range(3)]
    m = [[cq.X(q[0], "alpha"), cq.Y(q[1], "alpha"), cq.Z(q[2], "beta", scalar=0.5)],
         [cq.XX(q[0], q[1], "alpha", scalar=2.0), cq.H(q[2])]]
    prog = cq.serialize(m)
    other = cq.serialize([[cq.X(q[0], "beta")]])
    out = ops.tfq_ps_symbol_replace([prog, other], ["alpha", "beta", "gamma"],
                                    ["new", "old", "unused"])
    assert out.shape == (2, 3, 3)
    base = _ops_of(prog)
    # alpha occurs three times in program 0
    for k in range(3):
        got = _ops_of(out[0, 0, k])
        diff = [(a, b) for a, b in zip(base, got) if a != b]
        assert len(diff) == 1
        a, b = diff[0]
        assert a[2]["exponent"] == "sym:alpha" and b[2]["exponent"] == "sym:new"
        assert {x: y for x, y in a[2].items() if x != "exponent"} == \
               {x: y for x, y in b[2].items() if x != "exponent"}
    assert [o[2]["exponent"] for o in _ops_of(out[0, 0, 0])][:2] == ["sym:new", "sym:alpha"]
    # beta once, then padding; gamma never
    assert _ops_of(out[0, 1, 0])[2][2]["exponent"] == "sym:old"
    for e in (out[0, 1, 1], out[0, 1, 2], out[0, 2, 0], out[1, 0, 0], out[1, 2, 2]):
        p = _parse(e)
        assert p.language.gate_set == "tfq_gate_set" and len(p.circuit.moments) == 0
        assert p.WhichOneof("program") == "circuit"
    assert _ops_of(out[1, 1, 0])[0][2]["exponent"] == "sym:old"
    # the replaced programs still run through the simulation ops' parser
    d = ops.host_describe_plan(out[0, 0, 1], ["alpha", "new", "beta"])
    assert d["n"] == 3


def test_symbol_replace_errors():
    """:400-435"""
    q = cq.grid(0, 0)
    prog = cq.serialize([[cq.X(q, "alpha")]])
    with pytest.raises(ops.InvalidArgumentError, match="symbols.shape is not equal"):
        ops.tfq_ps_symbol_replace([prog], ["alpha"], ["a", "b"])
    with pytest.raises(ops.InvalidArgumentError, match="rank 1"):
        ops.tfq_ps_symbol_replace([[prog]], ["alpha"], ["a"])
    with pytest.raises(ops.InvalidArgumentError, match="rank 1"):
        ops.tfq_ps_symbol_replace([prog], [["alpha"]], ["a"])
    with pytest.raises(ops.InvalidArgumentError, match="Unparseable proto"):
        ops.tfq_ps_symbol_replace([b"\xff\xfejunk"], ["alpha"], ["a"])


# ------------------------------------------------------- weights from symbols
def test_weights_from_symbols():
    """:674-853: exponent_scalar per symbol occurrence, zero padded, symbol
    order given by `symbols`, composite / noise gates ignored."""
    q = [cq.grid(0, i) for i in range(3)]
    prog = cq.serialize([[cq.X(q[0], "alpha", scalar=5.0)]])
    np.testing.assert_array_equal(ops.tfq_ps_weights_from_symbols([prog], ["alpha"]),
                                  [[[5.0]]])
    # nothing parameterised: an empty last dimension
    none = cq.serialize([[cq.X(q[0], 0.5)]])
    assert ops.tfq_ps_weights_from_symbols([none], []).shape == (1, 0, 0)
    assert ops.tfq_ps_weights_from_symbols([none], ["alpha"]).shape == (1, 1, 0)
    # many values, out of order, padding
    m = [[cq.X(q[0], "a", scalar=2.0), cq.Y(q[1], "b", scalar=3.0), cq.Z(q[2], "a", scalar=4.0)],
         [cq.ZZ(q[0], q[1], "a", scalar=-1.5), cq.H(q[2])]]
    w = ops.tfq_ps_weights_from_symbols([cq.serialize(m), prog], ["b", "a", "alpha"])
    np.testing.assert_array_equal(
        w, [[[3.0, 0, 0], [2.0, 4.0, -1.5], [0, 0, 0]], [[0, 0, 0], [0, 0, 0], [5.0, 0, 0]]])
    assert w.dtype == np.float32
    # ignored gate ids (tfq_ps_weights_from_symbols_op.cc:76-80)
    ig = [[cq.ISWAP(q[0], q[1], "a", scalar=9.0), cq.PhasedX(q[2], "a", "a")],
          [cq.FSim(q[0], q[1], "a", "a"), cq.X(q[2], "a", scalar=0.25)],
          [cq.depolarize(q[0], 0.1)]]
    np.testing.assert_array_equal(
        ops.tfq_ps_weights_from_symbols([cq.serialize(ig)], ["a"]), [[[0.25]]])


def test_weights_from_symbols_errors():
    """:702-727"""
    q = cq.grid(0, 0)
    prog = cq.serialize([[cq.X(q, "alpha")]])
    with pytest.raises(ops.InvalidArgumentError, match="sympy.Symbol not found"):
        ops.tfq_ps_weights_from_symbols([prog], ["beta"])
    with pytest.raises(ops.InvalidArgumentError, match="rank 1"):
        ops.tfq_ps_weights_from_symbols([[prog]], ["alpha"])
    with pytest.raises(ops.InvalidArgumentError, match="rank 1"):
        ops.tfq_ps_weights_from_symbols([prog], [["alpha"]])
    with pytest.raises(ops.InvalidArgumentError, match="Unparseable proto"):
        ops.tfq_ps_weights_from_symbols([b"\xff\xfejunk"], ["alpha"])


def test_weights_reference_vectors():
    """The reference's own expected tensors (tfq_ps_util_ops_test.py:729-783:
    test_many_values, test_many_symbols, test_out_of_order)."""
    bit = cq.line(1)
    circuits = [
        cq.serialize([[cq.X(bit, "alpha", scalar=2.0)], [cq.Y(bit, "alpha", scalar=3.0)],
                      [cq.Z(bit, "alpha")], [cq.X(bit, "alpha", scalar=4.0)]]),
        cq.serialize([[cq.X(bit, "alpha", scalar=9.0)]]),
        cq.serialize([[cq.X(bit, "beta")]]),
    ]
    np.testing.assert_allclose(
        ops.tfq_ps_weights_from_symbols(circuits, ["alpha", "beta"]),
        np.array([[[2.0, 3.0, 1.0, 4.0], [0.0, 0.0, 0.0, 0.0]],
                  [[9.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]],
                  [[0.0, 0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0]]]))
    bit = cq.grid(0, 0)
    circuits = [cq.serialize([[cq.X(bit, s, scalar=c)]]) for s, c in
                (("alpha", 2.0), ("beta", 6.0), ("alpha", 5.0), ("gamma", 8.0), ("delta", 9.0))]
    np.testing.assert_allclose(
        ops.tfq_ps_weights_from_symbols(circuits, ["alpha", "beta", "gamma", "delta"]),
        np.array([[[2.0], [0.0], [0.0], [0.0]], [[0.0], [6.0], [0.0], [0.0]],
                  [[5.0], [0.0], [0.0], [0.0]], [[0.0], [0.0], [8.0], [0.0]],
                  [[0.0], [0.0], [0.0], [9.0]]]))
    c = cq.serialize([[cq.X(bit, "alpha", scalar=2.0)], [cq.Y(bit, "beta", scalar=3.0)]])
    np.testing.assert_allclose(ops.tfq_ps_weights_from_symbols([c], ["alpha", "beta"]),
                               [[[2.0], [3.0]]])
    np.testing.assert_allclose(ops.tfq_ps_weights_from_symbols([c], ["beta", "alpha"]),
                               [[[3.0], [2.0]]])


def test_symbol_replace_weight_coefficient():
    """tfq_ps_util_ops_test.py:437-478: scalar multiples survive the rename;
    checked like there through the unitary at alpha = 1.23, new = 4.56."""
    bit = cq.grid(0, 0)
    gates = (cq.X, cq.Y, cq.Z)
    scal = (2.4, 3.4, 4.4)
    prog = cq.serialize([[g(bit, "alpha", scalar=c)] for g, c in zip(gates, scal)])
    out = ops.tfq_ps_symbol_replace([prog], ["alpha"], ["new"])
    assert out.shape == (1, 1, 3)
    vals = np.array([1.23, 4.56])
    for i in range(3):
        want = cq.serialize([[g(bit, "new" if k == i else "alpha", scalar=c)]
                             for k, (g, c) in enumerate(zip(gates, scal))])
        np.testing.assert_allclose(_unitary(out[0, 0, i], ["alpha", "new"], vals),
                                   _unitary(want, ["alpha", "new"], vals), atol=1e-5)


def test_encoder_round_trip_on_random_circuits():
    """A program without parameterised composite gates passes through
    TfqPsDecompose unchanged: decode (wire.cc) + encode (ps_ops.cc) keeps every
    operation, argument (floats, symbols, control strings) and qubit id, for
    binary and text-format inputs."""
    qs = [cq.grid(0, i) for i in range(3)] + [cq.grid(1, i) for i in range(3)] + [cq.line(7)]
    for seed in range(6):
        m = cq.random_circuit(qs, 8, 100 + seed, controls=True, symbols=("a", "b"))
        # literal composite gates stay; drop the parameterised ones
        keep = []
        for mom in m:
            row = [op for op in mom
                   if not (op.gate in ("ISP", "PXP", "FSIM", "PISP") and
                           any(isinstance(v, str) for v in op.args.values()))]
            if row:
                keep.append(row)
        for ser in (cq.serialize, cq.serialize_text):
            out = ops.tfq_ps_decompose([ser(keep)])
            assert _ops_of(out[0]) == _ops_of(cq.serialize(keep)), seed
