"""Pins the CPU oracle to every golden vector the reference's own tests hold
for the hot path. Both oracle executors (numpy, C) are run."""
import numpy as np
import pytest

from oracle import tfq_oracle as orc
from quantum_b200 import circuits as cq
from quantum_b200 import proto as pb

BACKENDS = ["numpy", "c"]


def _state_circuit():
    # util_qsim_test.cc:55-61: X^0.25 on qsim qubit 1, CX(control qsim 1 ->
    # target qsim 0), Y^0.5 on qsim qubit 0.  qsim qubit k = proto id n-1-k.
    q0, q1 = "0", "1"   # proto ids: id 0 <-> qsim qubit 1
    return [[cq.X(q0, 0.25)], [cq.CNOT(q0, q1)], [cq.Y(q1, 0.5)]]


GOLD_STATE = np.array([0.25 + 0.60355j, 0.25 + 0.60355j, -0.25 + 0.10355j,
                       0.25 - 0.10355j])


@pytest.mark.parametrize("backend", BACKENDS)
def test_state_golden(backend):
    """util_qsim_test.cc:510-517 (amplitude index bit k = qsim qubit k)."""
    st = orc.simulate_state([cq.serialize(_state_circuit())], [], np.zeros(
        (1, 0), np.float32), backend=backend)
    np.testing.assert_allclose(st[0], GOLD_STATE, atol=1e-5)


TWO_TERM = [("ZZ", 0.0), ("ZX", 0.1234), ("ZY", 0.0), ("XZ", 0.0),
            ("XX", 0.0), ("XY", -0.08725), ("YZ", 0.08725), ("YX", 0.0),
            ("YY", 0.0)]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("pp,gold", TWO_TERM)
def test_two_term_expectation(backend, pp, gold):
    """util_qsim_test.cc:122-181."""
    ps = cq.pauli_sum([(0.1234, [("0", pp[0]), ("1", pp[1])])])
    out = orc.simulate_expectation([cq.serialize(_state_circuit())], [],
                                   np.zeros((1, 0), np.float32), [[ps]],
                                   backend=backend)
    assert abs(out[0, 0] - gold) < 1e-5


@pytest.mark.parametrize("pp,gold", TWO_TERM)
def test_two_term_sampled_expectation(pp, gold):
    """util_qsim_test.cc:49-116 (1e6 shots, tol 1e-2)."""
    ps = cq.pauli_sum([(0.1234, [("0", pp[0]), ("1", pp[1])])])
    out = orc.simulate_sampled_expectation(
        [cq.serialize(_state_circuit())], [], np.zeros((1, 0), np.float32),
        [[ps]], [[200000]], seed=1234)
    assert abs(out[0, 0] - gold) < 1e-2


COMPOUND = [(0.1234, [("0", "Z"), ("1", "X")]), (-3.0, [("0", "X")]),
            (4.0, [])]


@pytest.mark.parametrize("backend", BACKENDS)
def test_compound_expectation(backend):
    """util_qsim_test.cc:299-355: 0.1234 ZX - 3 X + 4 I -> 4.1234."""
    out = orc.simulate_expectation([cq.serialize(_state_circuit())], [],
                                   np.zeros((1, 0), np.float32),
                                   [[cq.pauli_sum(COMPOUND)]], backend=backend)
    assert abs(out[0, 0] - 4.1234) < 1e-5


def test_compound_sampled_expectation():
    """util_qsim_test.cc:236-297."""
    out = orc.simulate_sampled_expectation(
        [cq.serialize(_state_circuit())], [], np.zeros((1, 0), np.float32),
        [[cq.pauli_sum(COMPOUND)]], [[3000000]], seed=11)
    assert abs(out[0, 0] - 4.1234) < 1e-2


@pytest.mark.parametrize("backend", BACKENDS)
def test_accumulate_operators(backend):
    """util_qsim_test.cc:437-528."""
    prog = pb.Program()
    prog.ParseFromString(cq.serialize(_state_circuit()))
    s1, s2 = pb.PauliSum(), pb.PauliSum()
    s1.ParseFromString(cq.pauli_sum(COMPOUND))
    s2.ParseFromString(cq.pauli_sum([(-5.0, [])]))
    n = orc.resolve_qubit_ids(prog, [s1, s2])
    gates = orc.circuit_from_program(prog, {}, n)
    steps = orc.forward_steps(gates) + orc.accumulate_steps(
        [s1, s2], [0.5, 0.25], orc.SV, orc.SCRATCH, orc.SCRATCH2)
    if backend == "numpy":
        vm = orc.NumpyVM(n)
        vm.run(steps, 1)
        sv, scratch, dest = vm.bufs
    else:
        # C VM returns buffer 0 only: move dest/scratch there in two runs.
        _, dest = orc._run(steps + [("copy", orc.SCRATCH2, orc.SV)], n, 0,
                           "c", True)
        _, scratch = orc._run(steps + [("copy", orc.SCRATCH, orc.SV)], n, 0,
                              "c", True)
        _, sv = orc._run(steps, n, 0, "c", True)
    gold = np.array([0.577925 + 0.334574j, -0.172075 + 0.645234j,
                     -0.577925 - 0.821275j, -0.172075 - 0.989384j])
    np.testing.assert_allclose(dest, gold, atol=1e-5)
    np.testing.assert_allclose(sv, GOLD_STATE, atol=1e-5)
    np.testing.assert_allclose(scratch, GOLD_STATE, atol=1e-5)


@pytest.mark.parametrize("backend", BACKENDS)
def test_dagger_round_trip(backend):
    """util_qsim_test.cc:357-435: U then U^dagger returns to |00>."""
    prog = pb.Program()
    prog.ParseFromString(cq.serialize(_state_circuit()))
    n = orc.resolve_qubit_ids(prog)
    gates = orc.circuit_from_program(prog, {}, n)
    fused = orc.basic_fuse(gates)
    steps = orc.forward_steps(gates)
    for f in reversed(fused):
        steps.append(orc._apply_step(orc.SV, f, orc.dagger(f.matrix)))
    _, st = orc._run(steps, n, 0, backend, True)
    np.testing.assert_allclose(st, [1, 0, 0, 0], atol=1e-5)


def _flat(m):
    return np.asarray(m, np.complex64).reshape(-1).view(np.float32)


def _g(kind, qubits, params, symbols, placeholders):
    return orc.Gate(kind, qubits, None, (), (), tuple(np.float32(p) for p in
                                                       params),
                    symbols, placeholders)


def test_gradient_gate_goldens():
    """adj_util_test.cc:355-560 (8 / 32 floats each)."""
    g = _g("YP", (2,), (0.125, 1.0, 0.0), ["hello"], ["exponent"])
    np.testing.assert_allclose(
        _flat(orc.gradient_matrix(g, "exponent")),
        [-0.60111, 1.45122, -1.45122, -0.60111, 1.45122, 0.60111, -0.60111,
         1.45122], atol=1e-4)

    g = _g("XXP", (2, 3), (0.001, 1.0, 0.0), ["hi"], ["exponent"])
    a, b = -0.004934, 1.57078
    exp = np.zeros((4, 4), np.complex64)
    for i in range(4):
        exp[i, i] = a + 1j * b
        exp[i, 3 - i] = -a - 1j * b
    np.testing.assert_allclose(orc.gradient_matrix(g, "exponent"), exp,
                               atol=1e-4)

    g = _g("PXP", (2,), (10.123, 1.0, 1.0, 1.0, 0.0), ["h2"],
           ["phase_exponent"])
    np.testing.assert_allclose(
        _flat(orc.gradient_matrix(g, "phase_exponent")),
        [0, 0, -1.18397, -2.9099, -1.18397, 2.9099, 0, 0], atol=1e-3)

    g = _g("PXP", (2,), (10.123, 1.0, 0.789, 1.0, 0.0), ["h3"], ["exponent"])
    np.testing.assert_allclose(
        _flat(orc.gradient_matrix(g, "exponent")),
        [-0.96664, -1.23814, 1.36199, 0.78254, 0.42875, 1.51114, -0.96664,
         -1.23814], atol=1e-3)

    g = _g("FSIM", (2, 3), (0.5, 1.0, 1.2, 1.0), ["hihi"], ["theta"])
    exp = np.zeros((4, 4), np.complex64)
    exp[1, 1] = exp[2, 2] = -0.47942
    exp[1, 2] = exp[2, 1] = -0.87758j
    np.testing.assert_allclose(orc.gradient_matrix(g, "theta"), exp, atol=1e-4)

    exp = np.zeros((4, 4), np.complex64)
    exp[3, 3] = -0.932039 - 0.362357j
    np.testing.assert_allclose(orc.gradient_matrix(g, "phi"), exp, atol=1e-4)

    g = _g("PISP", (3, 2), (8.9, 1.0, -3.2, 1.0), ["h"], ["phase_exponent"])
    exp = np.zeros((4, 4), np.complex64)
    exp[1, 2] = -4.83441 + 3.51238j
    exp[2, 1] = 4.83441 + 3.51238j
    np.testing.assert_allclose(orc.gradient_matrix(g, "phase_exponent"), exp,
                               atol=1e-3)
    exp = np.zeros((4, 4), np.complex64)
    exp[1, 1] = exp[2, 2] = -1.49391
    exp[1, 2] = 0.285312 + 0.392698j
    exp[2, 1] = -0.285312 + 0.392698j
    np.testing.assert_allclose(orc.gradient_matrix(g, "exponent"), exp,
                               atol=1e-3)


def _adj(circuit, names, vals, backend):
    q0, q1 = cq.grid(0, 0), cq.grid(0, 1)
    ops = [cq.pauli_sum([(1.0, [(q0, "Z")])]), cq.pauli_sum([(1.0, [(q1, "X")])])]
    return orc.adjoint_gradient([cq.serialize(circuit)], names,
                                np.array([vals], np.float32), [ops],
                                np.ones((1, 2), np.float32), backend=backend)


@pytest.mark.parametrize("backend", BACKENDS)
def test_adjoint_goldens(backend):
    """tfq_adj_grad_op_test.py:265-397 (atol 1e-3 as in the reference)."""
    q0, q1 = cq.grid(0, 0), cq.grid(0, 1)
    base = [[cq.X(q0, "alpha"), cq.Y(q1, "beta")], [cq.CNOT(q0, q1)]]
    out = _adj(base, ["alpha", "beta"], [0.123, 0.456], backend)
    np.testing.assert_allclose(out, [[-1.18392, 0.43281]], atol=1e-3)

    c2 = base + [[cq.FSim(q0, q1, "gamma", 0.5)]]
    out = _adj(c2, ["alpha", "beta", "gamma"], [0.123, 0.456, 0.789], backend)
    np.testing.assert_allclose(out, [[-2.100, -1.7412, -1.5120]], atol=1e-3)

    c3 = base + [[cq.FSim(q0, q1, "gamma", "gamma")]]
    out = _adj(c3, ["alpha", "beta", "gamma"], [0.123, 0.456, 0.789], backend)
    np.testing.assert_allclose(out, [[-2.3484, -1.7532, -1.64264]], atol=1e-3)

    l0, l1 = cq.line(0), cq.line(1)
    c4 = [[cq.X(l0, "alpha"), cq.Y(l1, "alpha")], [cq.CNOT(l0, l1)],
          [cq.FSim(l0, l1, -0.56, "alpha")]]
    ops = [cq.pauli_sum([(1.0, [(l0, "Z")])]), cq.pauli_sum([(1.0, [(l1, "X")])])]
    out = orc.adjoint_gradient([cq.serialize(c4)], ["alpha", "beta", "gamma"],
                               np.array([[0.123, 0.456, 0.789]], np.float32),
                               [ops], np.ones((1, 2), np.float32),
                               backend=backend)
    np.testing.assert_allclose(out, [[1.2993, 0, 0]], atol=1e-3)


@pytest.mark.parametrize("backend", BACKENDS)
def test_docstring_and_gradient_test_goldens(backend):
    """circuit_execution_ops.py:52-68: H^0.123 with 3.5 X - 2.2 Y ->
    0.71530885; gradient_test.py:230-248: d<Z>/da of X^a at 0.123 ->
    -1.1839752 (tol 1e-2)."""
    q = cq.grid(0, 0)
    ps = cq.pauli_sum([(3.5, [(q, "X")]), (-2.2, [(q, "Y")])])
    out = orc.simulate_expectation([cq.serialize([[cq.H(q, "alpha")]])],
                                   ["alpha"], np.array([[0.123]], np.float32),
                                   [[ps]], backend=backend)
    assert abs(out[0, 0] - 0.71530885) < 1e-5
    z = cq.pauli_sum([(1.0, [(q, "Z")])])
    g = orc.adjoint_gradient([cq.serialize([[cq.X(q, "alpha")]])], ["alpha"],
                             np.array([[0.123]], np.float32), [[z]],
                             np.ones((1, 1), np.float32), backend=backend)
    assert abs(g[0, 0] - (-1.1839752)) < 1e-2


def test_samples_padding_golden():
    """tfq_simulate_ops_test.py:319-346: X on all qubits -> all ones;
    shorter circuits are left-padded with -2."""
    progs = []
    for n in (3, 5):
        progs.append(cq.serialize([[cq.X(cq.grid(0, i)) for i in range(n)]]))
    out = orc.simulate_samples(progs, [], np.zeros((2, 0), np.float32), [4])
    assert out.shape == (2, 4, 5)
    assert (out[1] == 1).all()
    assert (out[0, :, :2] == -2).all() and (out[0, :, 2:] == 1).all()


def test_state_padding_golden():
    """tfq_simulate_ops_test.py:477-492 and empty circuit -> [1,-2,-2,..]."""
    progs = [cq.serialize([[cq.X(cq.grid(0, 0))]]),
             cq.serialize([[cq.X(cq.grid(0, i)) for i in range(2)]]),
             cq.serialize([])]
    out = orc.simulate_state(progs, [], np.zeros((3, 0), np.float32))
    np.testing.assert_allclose(out[0], [0, 1, -2, -2], atol=1e-6)
    np.testing.assert_allclose(out[1], [0, 0, 0, 1], atol=1e-6)
    np.testing.assert_allclose(out[2], [1, -2, -2, -2], atol=1e-6)


def test_resolve_qubit_ids():
    """program_resolution_test.cc:184-290: grid ids sort by (row, col), line
    qubits after all grid qubits, control qubits are registered too."""
    c = [[cq.X("0_1"), cq.X("0_0").controlled_by(["1_0", "5"], [1, 0])]]
    prog = pb.Program()
    prog.ParseFromString(cq.serialize(c))
    ps = pb.PauliSum()
    ps.ParseFromString(cq.pauli_sum([(1.0, [("5", "Z"), ("0_0", "X")])]))
    n = orc.resolve_qubit_ids(prog, [ps])
    assert n == 4
    ops = prog.circuit.moments[0].operations
    assert ops[0].qubits[0].id == "1"
    assert ops[1].qubits[0].id == "0"
    assert ops[1].args["control_qubits"].arg_value.string_value == "2,3"
    assert [p.qubit_id for p in ps.terms[0].paulis] == ["3", "0"]
    bad = pb.PauliSum()
    bad.ParseFromString(cq.pauli_sum([(1.0, [("9_9", "Z")])]))
    with pytest.raises(orc.InvalidArgumentError, match="qubits not found"):
        orc.resolve_qubit_ids(pb.Program.FromString(cq.serialize(c)), [bad])


def test_gate_closed_forms_are_unitary_and_consistent():
    """Unpinned-by-golden closed forms: unitarity + Cirq identities."""
    rng = np.random.default_rng(0)
    for name, fn in list(orc.EIGEN_1Q.items()) + list(orc.EIGEN_2Q.items()):
        for _ in range(4):
            t, s = rng.uniform(-2, 2), rng.uniform(-1, 1)
            u = fn(t, s).astype(np.complex128)
            np.testing.assert_allclose(u @ u.conj().T, np.eye(len(u)),
                                       atol=2e-6)
            # U(t)^2 == U(2t)
            np.testing.assert_allclose(u @ u, fn(2 * t, s), atol=5e-6)
    X, Y, Z = orc.mat_xpow(1), orc.mat_ypow(1), orc.mat_zpow(1)
    np.testing.assert_allclose(X, [[0, 1], [1, 0]], atol=1e-6)
    np.testing.assert_allclose(Y, [[0, -1j], [1j, 0]], atol=1e-6)
    np.testing.assert_allclose(Z, [[1, 0], [0, -1]], atol=1e-6)
    np.testing.assert_allclose(orc.mat_hpow(1), np.array([[1, 1], [1, -1]]) /
                               np.sqrt(2), atol=1e-6)
    np.testing.assert_allclose(orc.mat_xxpow(1), np.kron(X, X), atol=1e-6)
    np.testing.assert_allclose(orc.mat_yypow(1), np.kron(Y, Y), atol=1e-6)
    np.testing.assert_allclose(orc.mat_zzpow(1), np.kron(Z, Z), atol=1e-6)
    np.testing.assert_allclose(orc.mat_czpow(1), np.diag([1, 1, 1, -1]),
                               atol=1e-6)
    cn = np.eye(4)[[0, 1, 3, 2]]
    np.testing.assert_allclose(orc.mat_cxpow(1), cn, atol=1e-6)
    sw = np.eye(4)[[0, 2, 1, 3]]
    np.testing.assert_allclose(orc.mat_swappow(1), sw, atol=1e-6)
    isw = np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]])
    np.testing.assert_allclose(orc.mat_iswappow(1), isw, atol=1e-6)
    # PhasedX(p, t) = Z^p X^t Z^-p ; PhasedISwap(p=0) = ISWAP^t
    p, t = 0.37, 0.81
    np.testing.assert_allclose(
        orc.mat_phasedxpow(p, t),
        orc.mat_zpow(p) @ orc.mat_xpow(t) @ orc.mat_zpow(-p), atol=2e-6)
    np.testing.assert_allclose(orc.mat_phasediswappow(0.0, t),
                               orc.mat_iswappow(t), atol=2e-6)
    # FSim(theta, 0) = ISWAP^(-2 theta / pi)
    th = 0.3
    np.testing.assert_allclose(orc.mat_fsim(th, 0.0),
                               orc.mat_iswappow(-2 * th / np.pi), atol=2e-6)


def test_c_and_numpy_executors_agree_on_random_circuits():
    qs = [cq.grid(0, i) for i in range(5)] + [cq.line(2)]
    progs, sums = [], []
    for s in range(4):
        progs.append(cq.serialize(cq.random_circuit(qs, 8, seed=s,
                                                    controls=True)))
        sums.append([cq.random_pauli_sum(qs, 4, seed=100 + s),
                     cq.random_pauli_sum(qs, 4, seed=200 + s)])
    vals = np.zeros((4, 0), np.float32)
    a = orc.simulate_expectation(progs, [], vals, sums, backend="numpy")
    b = orc.simulate_expectation(progs, [], vals, sums, backend="c", threads=2)
    np.testing.assert_allclose(a, b, atol=2e-6)
    sa = orc.simulate_state(progs, [], vals, backend="numpy")
    sb = orc.simulate_state(progs, [], vals, backend="c")
    np.testing.assert_allclose(sa, sb, atol=2e-6)
    # fused and unfused sweeps agree
    prog = pb.Program.FromString(progs[0])
    n = orc.resolve_qubit_ids(prog)
    gates = orc.circuit_from_program(prog, {}, n)
    _, f = orc._run(orc.forward_steps(gates, True), n, 0, "numpy", True)
    _, u = orc._run(orc.forward_steps(gates, False), n, 0, "numpy", True)
    np.testing.assert_allclose(f, u, atol=2e-6)
    assert abs(np.vdot(f, f) - 1) < 1e-5


def test_sample_tree_matches_sequential_walk_and_philox_is_uniform():
    rng = np.random.default_rng(3)
    st = (rng.normal(size=64) + 1j * rng.normal(size=64)).astype(np.complex64)
    st /= np.linalg.norm(st)
    u = np.sort(rng.random(2000))
    idx = orc.sample_tree(st, u)
    p = (st.real.astype(np.float64) ** 2 + st.imag.astype(np.float64) ** 2)
    cs = np.cumsum(p)
    ref = np.searchsorted(cs, u * cs[-1], side="right")
    assert (idx == np.minimum(ref, 63)).mean() > 0.999
    assert (np.diff(idx) >= 0).all()
    ph = orc.philox_uniforms(42, 3, 1, 2, 100000)
    assert 0 <= ph.min() and ph.max() < 1
    assert abs(ph.mean() - 0.5) < 5e-3
    assert len(np.unique(ph)) == len(ph)
    # known-answer: Philox4x32-10 with zero key/counter (Random123 KAT)
    z = orc.philox_uniforms(0, 0, 0, 0, 1)
    x = int(z[0] * 2 ** 53)
    assert x == ((0x6627e8d5 << 32 | 0xe169c58d) >> 11)


def test_committed_fixtures_match_the_oracle():
    """tests/golden/*.npz (made by tests/golden/make_golden.py) still equal
    what the oracle computes: a drift of the oracle shows up here first."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(here, "five_ops_ragged.npz"), allow_pickle=True)
    progs = list(z["programs"])
    names = list(z["symbol_names"])
    sums = [list(r) for r in z["pauli_sums"]]
    vals = z["symbol_values"]
    np.testing.assert_allclose(orc.simulate_state(progs, names, vals), z["state"],
                               atol=1e-7)
    np.testing.assert_allclose(orc.simulate_expectation(progs, names, vals, sums),
                               z["expectation"], atol=1e-6)
    np.testing.assert_allclose(
        orc.adjoint_gradient(progs, names, vals, sums, z["downstream"]),
        z["gradient"], atol=1e-5)
    np.testing.assert_array_equal(
        orc.simulate_samples(progs, names, vals, [64], uniforms=z["uniforms"]),
        z["samples"])


def test_inner_product_grad_oracle_is_the_derivative_of_the_inner_product():
    """No offline golden exists for TfqInnerProductGrad (the reference test
    compares with a live cirq): the restatement is checked against central
    differences of the (pinned) forward inner product, real and imaginary
    part, including a symbol shared by several gates, a controlled
    parameterised gate and the downstream weights."""
    q = [cq.grid(0, i) for i in range(4)]
    circ = [[cq.H(x) for x in q],
            [cq.X(q[0], "a"), cq.Y(q[1], "a"), cq.ZZ(q[2], q[3], "b")],
            [cq.CNOT(q[0], q[3]), cq.FSim(q[1], q[2], 0.3, "b")]]
    prog = cq.serialize(circ)
    others = [cq.serialize(cq.random_circuit(q, 5, s) + [[cq.H(x) for x in q]])
              for s in (3, 4)]
    v = np.array([[0.37, 1.21]], np.float32)
    down = np.array([[0.7, -1.3]], np.float32)
    g = orc.inner_product_grad([prog], ["a", "b"], v, [others], down)
    assert g.shape == (1, 2) and g.dtype == np.complex64
    for col in range(2):
        dv = np.zeros_like(v)
        dv[0, col] = 5e-3
        ip_p = orc.inner_product([prog], ["a", "b"], v + dv, [others])[0]
        ip_m = orc.inner_product([prog], ["a", "b"], v - dv, [others])[0]
        fd = np.sum(down[0] * (ip_p - ip_m)) / 1e-2
        assert abs(g[0, col] - fd) < 2e-3
    with pytest.raises(orc.InvalidArgumentError, match="positive integer"):
        orc.inner_product_grad([prog], [], np.zeros((1, 0), np.float32), [others], down)
