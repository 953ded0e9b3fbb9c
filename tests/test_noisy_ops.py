"""Next-row N2: the noisy trajectory ops TfqNoisyExpectation,
TfqNoisySampledExpectation, TfqNoisySamples
(tensorflow_quantum/core/ops/noise/*.cc).

CPU: the oracle's trajectory restatement (qsim's channel definitions and
Kraus sampling, which are NOT under /root/reference: parity unpinned, see
oracle/tfq_oracle.py section 6) converges to the exact density-matrix
evolution of the same channels -- the mathematical definition cirq gives them.
GPU: the CUDA path against the oracle with identical uniforms (the draws are
then the same Kraus operators, so values agree to float32 round-off), and the
device Philox streams against the oracle's.
"""
import numpy as np
import pytest

from oracle import tfq_oracle as orc
from quantum_b200 import circuits as cq
from quantum_b200 import ops


def _noisy_circuit(q, sym="a"):
    return [
        [cq.H(q[0]), cq.X(q[1], 0.3), cq.Y(q[2], sym)],
        [cq.depolarize(q[0], 0.2), cq.amplitude_damp(q[1], 0.3)],
        [cq.CNOT(q[0], q[1]), cq.Z(q[2], 0.4)],
        [cq.generalized_amplitude_damp(q[2], 0.4, 0.2), cq.phase_damp(q[1], 0.5)],
        [cq.ISWAP(q[1], q[2], 0.5), cq.bit_flip(q[0], 0.1)],
        [cq.phase_flip(q[0], 0.2), cq.asymmetric_depolarize(q[1], 0.05, 0.1, 0.15),
         cq.reset(q[2])],
        [cq.X(q[2], 0.7), cq.H(q[0])],
    ]


def _density_matrix(prog, smap, n):
    p = orc.parse_proto(prog, orc._pb.Program)
    orc.resolve_qubit_ids(p)
    items = orc.noisy_circuit_from_program(p, smap, n)
    dim = 2 ** n
    rho = np.zeros((dim, dim), complex)
    rho[0, 0] = 1

    def embed(m, axis):
        mats = [np.eye(2)] * n
        mats[axis] = m
        out = mats[0]
        for x in mats[1:]:
            out = np.kron(out, x)
        return out

    for it in items:
        if it[0] == "gate":
            g = it[1]
            U = np.zeros((dim, dim), complex)
            for b in range(dim):
                v = np.zeros(dim, np.complex64)
                v[b] = 1
                orc._np_apply(v, n, g.qubits, g.matrix, g.controls, g.cvalues)
                U[:, b] = v
            rho = U @ rho @ U.conj().T
        else:
            _, ax, kraus = it
            new = np.zeros_like(rho)
            for unitary, prob, m in kraus:
                K = embed(m.astype(complex), ax) * (np.sqrt(prob) if unitary else 1.0)
                new += K @ rho @ K.conj().T
            rho = new
    return rho, embed


def test_oracle_trajectories_converge_to_density_matrix():
    q = [cq.grid(0, i) for i in range(3)]
    prog = cq.serialize(_noisy_circuit(q))
    rho, embed = _density_matrix(prog, {"a": (0, np.float32(0.7))}, 3)
    assert abs(np.trace(rho).real - 1) < 1e-6          # the channels are trace preserving
    X = np.array([[0, 1], [1, 0]])
    Y = np.array([[0, -1j], [1j, 0]])
    Z = np.diag([1.0, -1.0])
    O1 = embed(Z, 0) + 0.5 * embed(X, 1) @ embed(Y, 2)
    O2 = embed(Z, 2) - embed(Z, 1)
    exact = [np.trace(rho @ O1).real, np.trace(rho @ O2).real]
    sums = [[cq.pauli_sum([(1.0, [(q[0], "Z")]), (0.5, [(q[1], "X"), (q[2], "Y")])]),
             cq.pauli_sum([(1.0, [(q[2], "Z")]), (-1.0, [(q[1], "Z")])])]]
    T = 1500
    e = orc.noisy_expectation([prog], ["a"], np.array([[0.7]], np.float32), sums,
                              [[T, T]], seed=11)
    # per-trajectory values are bounded by 1.5 / 2: 4 sigma of the mean
    assert abs(e[0, 0] - exact[0]) < 4 * 1.5 / np.sqrt(T)
    assert abs(e[0, 1] - exact[1]) < 4 * 2.0 / np.sqrt(T)
    s = orc.noisy_samples([prog], ["a"], np.array([[0.7]], np.float32), [T], seed=5)
    p1 = [(1 - np.trace(rho @ embed(Z, k)).real) / 2 for k in range(3)]
    assert np.abs(s[0].mean(axis=0) - p1).max() < 4 * 0.5 / np.sqrt(T)
    se = orc.noisy_sampled_expectation([prog], ["a"], np.array([[0.7]], np.float32), sums,
                                       [[T, T]], seed=9)
    assert abs(se[0, 0] - exact[0]) < 5 * 1.5 / np.sqrt(T)
    assert abs(se[0, 1] - exact[1]) < 5 * 2.0 / np.sqrt(T)


def test_noiseless_program_through_noisy_oracle_equals_plain_expectation():
    q = [cq.grid(0, i) for i in range(4)]
    m = cq.random_circuit(q, 5, 3, symbols=("a",))
    prog = cq.serialize(m)
    sums = [[cq.random_pauli_sum(q, 4, 2, max_weight=3)]]
    v = np.array([[0.3]], np.float32)
    a = orc.noisy_expectation([prog], ["a"], v, sums, [[2]])
    b = orc.simulate_expectation([prog], ["a"], v, sums)
    np.testing.assert_allclose(a, b, atol=2e-6)


def _batch():
    q3 = [cq.grid(0, i) for i in range(3)]
    q5 = [cq.grid(1, i) for i in range(5)]
    m5 = cq.random_circuit(q5, 6, 21, symbols=("a",))
    m5.insert(2, [cq.depolarize(q5[0], 0.3), cq.amplitude_damp(q5[3], 0.4)])
    m5.insert(5, [cq.generalized_amplitude_damp(q5[1], 0.6, 0.3), cq.reset(q5[4]),
                  cq.phase_damp(q5[2], 0.2)])
    m5.append([cq.bit_flip(q5[2], 0.2), cq.phase_flip(q5[0], 0.25)])
    progs = [cq.serialize(_noisy_circuit(q3)), cq.serialize(m5), cq.serialize([]),
             cq.serialize(_noisy_circuit(q3, "b"))]
    names = ["a", "b"]
    vals = np.array([[0.7, 0.1], [0.2, 0.9], [0.0, 0.0], [0.5, 1.3]], np.float32)
    sums = [[cq.pauli_sum([(1.0, [(q3[0], "Z")]), (0.5, [(q3[1], "X"), (q3[2], "Y")])]),
             cq.pauli_sum([(1.0, [(q3[2], "Z")]), (0.25, [])])],
            [cq.random_pauli_sum(q5, 4, 8, max_weight=3),
             cq.pauli_sum([(1.0, [(x, "Z")]) for x in q5])],
            [cq.pauli_sum([(1.0, [(q3[0], "Z")])]), cq.pauli_sum([(1.0, [(q3[0], "Z")])])],
            [cq.pauli_sum([(1.0, [(q3[1], "Y")])]), cq.pauli_sum([(0.5, [(q3[0], "X")])])]]
    return progs, names, vals, sums


@pytest.mark.gpu
def test_noisy_expectation_matches_oracle():
    progs, names, vals, sums = _batch()
    ns = np.array([[40, 25], [30, 30], [3, 2], [17, 33]], np.int32)
    T, C = 40, 9
    u = np.random.default_rng(3).random((len(progs), T, C)).astype(np.float32)
    a = ops.tfq_noisy_expectation(progs, names, vals, sums, ns, uniforms=u)
    b = orc.noisy_expectation(progs, names, vals, sums, ns, uniforms=u)
    assert (a[2] == -2).all() and (b[2] == -2).all()
    np.testing.assert_allclose(a, b, atol=2e-5, rtol=1e-4)
    # the device Philox streams are the oracle's
    a = ops.tfq_noisy_expectation(progs, names, vals, sums, ns, seed=1234)
    b = orc.noisy_expectation(progs, names, vals, sums, ns, seed=1234)
    np.testing.assert_allclose(a, b, atol=2e-5, rtol=1e-4)


@pytest.mark.gpu
def test_noisy_samples_and_sampled_expectation_match_oracle():
    progs, names, vals, sums = _batch()
    S = 60
    a = ops.tfq_noisy_samples(progs, names, vals, [S], seed=77)
    b = orc.noisy_samples(progs, names, vals, [S], seed=77)
    assert a.shape == b.shape == (4, S, 5)
    assert (a[2] == -2).all() and (a[0, :, :2] == -2).all()
    # identical draws; a shot can only differ where float32 round-off moves a
    # CDF boundary across the measurement uniform
    assert (a != b).any(axis=2).mean() < 0.02
    ns = np.array([[50, 20], [30, 30], [2, 2], [40, 10]], np.int32)
    sa = ops.tfq_noisy_sampled_expectation(progs, names, vals, sums, ns, seed=5)
    sb = orc.noisy_sampled_expectation(progs, names, vals, sums, ns, seed=5)
    assert (sa[2] == -2).all()
    # one boundary crossing moves one term of one trajectory by 2 c / num_samples
    assert np.abs(sa - sb).max() < 0.15
    assert (np.abs(sa - sb) > 1e-5).mean() <= 0.5


@pytest.mark.gpu
def test_noisy_ops_errors_and_statistics():
    q = [cq.grid(0, i) for i in range(2)]
    prog = cq.serialize([[cq.X(q[0])], [cq.bit_flip(q[0], 0.25), cq.depolarize(q[1], 0.3)]])
    z0 = [[cq.pauli_sum([(1.0, [(q[0], "Z")])]), cq.pauli_sum([(1.0, [(q[1], "Z")])])]]
    v = np.zeros((1, 0), np.float32)
    e = ops.tfq_noisy_expectation([prog], [], v, z0, [[4000, 4000]], seed=3)
    # <Z0> = -(1 - 2 * 0.25), <Z1> = 1 - 2 * (2/3) * 0.3
    assert abs(e[0, 0] + 0.5) < 0.06 and abs(e[0, 1] - 0.6) < 0.06
    E = ops.InvalidArgumentError
    with pytest.raises(E, match="cirq.Channel"):        # plain ops still refuse channels
        ops.tfq_simulate_expectation([prog], [], v, z0)
    bad = cq.to_program([[cq.X(q[0])]])
    bad.circuit.moments[0].operations[0].gate.id = "QQ"
    with pytest.raises(E, match="Could not parse channel id: QQ"):
        ops.tfq_noisy_expectation([bad.SerializeToString()], [], v, z0, [[2, 2]])
    with pytest.raises(E, match="greater than 0"):
        ops.tfq_noisy_expectation([prog], [], v, z0, [[2, 0]])
    with pytest.raises(E, match="Dimension 1 of num_samples"):
        ops.tfq_noisy_expectation([prog], [], v, z0, [[2]])


# ---------------------------------------------------------------- N4: unitary
def test_oracle_unitary_is_unitary_and_matches_state():
    q = [cq.grid(0, i) for i in range(4)]
    prog = cq.serialize(cq.random_circuit(q, 6, 5, symbols=("a",)))
    v = np.array([[0.4]], np.float32)
    u = orc.calculate_unitary([prog], ["a"], v)[0]
    np.testing.assert_allclose(u @ u.conj().T, np.eye(16), atol=2e-6)
    np.testing.assert_allclose(u[:, 0], orc.simulate_state([prog], ["a"], v)[0], atol=1e-6)


@pytest.mark.gpu
def test_calculate_unitary_matches_oracle():
    """tfq_unitary_op_test.py: ragged batch, (-2, 0) padding, empty program."""
    progs, names, qss = [], ["a", "b"], []
    for k, n in enumerate([3, 5, 1, 6]):
        q = [cq.grid(0, i) for i in range(n)]
        progs.append(cq.serialize(cq.random_circuit(q, 6, 60 + k, symbols=("a", "b"))))
    progs.append(cq.serialize([]))
    vals = np.random.default_rng(1).uniform(0, 2, (5, 2)).astype(np.float32)
    a = ops.tfq_calculate_unitary(progs, names, vals)
    b = orc.calculate_unitary(progs, names, vals)
    assert a.shape == b.shape == (5, 64, 64)
    np.testing.assert_allclose(a, b, atol=2e-6)
    assert a[4, 0, 0] == 1 and (a[4].reshape(-1)[1:] == -2).all()
    with pytest.raises(ops.InvalidArgumentError, match="Number of circuits and values do not match"):
        ops.tfq_calculate_unitary(progs, names, vals[:2])
