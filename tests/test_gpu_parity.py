"""GPU parity: the CUDA path, called through the C ABI (quantum_b200.ops ->
libtfqb.so), against the CPU oracle on the same seeded inputs, and against
the reference's own golden vectors.

Tolerance (BASELINE.json north_star): expectations and gradients within
1e-5 absolute / 1e-4 relative in complex64; sampled bitstrings bit-exact for
identical uniforms.  State amplitudes: 2e-6 absolute (float32 round-off of a
different fusion order).
"""
import numpy as np
import pytest

from oracle import tfq_oracle as orc
from quantum_b200 import circuits as cq
from quantum_b200 import ops

pytestmark = pytest.mark.gpu

ATOL, RTOL = 1e-5, 1e-4


def _batch(n_list, seed, n_moments=8, controls=True, symbols=("a", "b", "c")):
    progs, qss = [], []
    for k, n in enumerate(n_list):
        qs = [cq.grid(0, i) for i in range(n)]
        if k % 2:  # mix in line qubits (sort after grid qubits)
            qs = [cq.grid(1, i) for i in range(n // 2)] + \
                [cq.line(i) for i in range(n - n // 2)]
        m = cq.random_circuit(qs, n_moments, seed + k, controls=controls,
                              symbols=symbols)
        progs.append(cq.serialize(m))
        qss.append(qs)
    return progs, qss


def _sums(qss, m_ops, seed, max_terms=5):
    return [[cq.random_pauli_sum(qs, max_terms, seed + 17 * i + j,
                                 max_weight=4) for j in range(m_ops)]
            for i, qs in enumerate(qss)]


# ------------------------------------------------------------------ goldens
def test_state_golden_util_qsim_test():
    """util_qsim_test.cc:437-464,510-517."""
    q0, q1 = cq.grid(0, 0), cq.grid(0, 1)
    # qsim qubit k <-> index bit k; proto qubit id i -> bit n-1-i
    circuit = [[cq.X(q0, 0.25)], [cq.CNOT(q0, q1)], [cq.Y(q1, 0.5)]]
    out = ops.tfq_simulate_state([cq.serialize(circuit)], [],
                                 np.zeros((1, 0), np.float32))
    ref = orc.simulate_state([cq.serialize(circuit)], [],
                             np.zeros((1, 0), np.float32))
    np.testing.assert_allclose(out, ref, atol=2e-6)


def test_adjoint_goldens_reference_op_test():
    """tfq_adj_grad_op_test.py:265-397 (reference atol 1e-3)."""
    q0, q1 = cq.grid(0, 0), cq.grid(0, 1)
    obs = [[cq.pauli_sum([(1.0, [(q0, "Z")])]),
            cq.pauli_sum([(1.0, [(q1, "X")])])]]
    ones = np.ones((1, 2), np.float32)
    base = [[cq.X(q0, "alpha"), cq.Y(q1, "beta")], [cq.CNOT(q0, q1)]]
    out = ops.tfq_adj_grad([cq.serialize(base)], ["alpha", "beta"],
                           [[0.123, 0.456]], obs, ones)
    np.testing.assert_allclose(out, [[-1.18392, 0.43281]], atol=1e-3)
    c2 = base + [[cq.FSim(q0, q1, "gamma", 0.5)]]
    out = ops.tfq_adj_grad([cq.serialize(c2)], ["alpha", "beta", "gamma"],
                           [[0.123, 0.456, 0.789]], obs, ones)
    np.testing.assert_allclose(out, [[-2.100, -1.7412, -1.5120]], atol=1e-3)
    c3 = base + [[cq.FSim(q0, q1, "gamma", "gamma")]]
    out = ops.tfq_adj_grad([cq.serialize(c3)], ["alpha", "beta", "gamma"],
                           [[0.123, 0.456, 0.789]], obs, ones)
    np.testing.assert_allclose(out, [[-2.3484, -1.7532, -1.64264]], atol=1e-3)
    l0, l1 = cq.line(0), cq.line(1)
    c4 = [[cq.X(l0, "alpha"), cq.Y(l1, "alpha")], [cq.CNOT(l0, l1)],
          [cq.FSim(l0, l1, -0.56, "alpha")]]
    obs_l = [[cq.pauli_sum([(1.0, [(l0, "Z")])]),
              cq.pauli_sum([(1.0, [(l1, "X")])])]]
    out = ops.tfq_adj_grad([cq.serialize(c4)], ["alpha", "beta", "gamma"],
                           [[0.123, 0.456, 0.789]], obs_l, ones)
    np.testing.assert_allclose(out, [[1.2993, 0, 0]], atol=1e-3)


def test_expectation_docstring_golden():
    """circuit_execution_ops.py:52-68 -> 0.71530885."""
    q = cq.grid(0, 0)
    ps = cq.pauli_sum([(3.5, [(q, "X")]), (-2.2, [(q, "Y")])])
    out = ops.tfq_simulate_expectation([cq.serialize([[cq.H(q, "alpha")]])],
                                       ["alpha"], [[0.123]], [[ps]])
    assert abs(out[0, 0] - 0.71530885) < 1e-5


def test_compound_expectation_golden():
    """util_qsim_test.cc:299-355: 0.1234 ZX - 3 X + 4 I -> 4.1234."""
    q0, q1 = cq.grid(0, 0), cq.grid(0, 1)
    circuit = [[cq.X(q0, 0.25)], [cq.CNOT(q0, q1)], [cq.Y(q1, 0.5)]]
    ps = cq.pauli_sum([(0.1234, [(q0, "Z"), (q1, "X")]), (-3.0, [(q0, "X")]),
                       (4.0, [])])
    a = ops.tfq_simulate_expectation([cq.serialize(circuit)], [],
                                     np.zeros((1, 0), np.float32), [[ps]])
    b = orc.simulate_expectation([cq.serialize(circuit)], [],
                                 np.zeros((1, 0), np.float32), [[ps]])
    np.testing.assert_allclose(a, b, atol=ATOL)


# ------------------------------------------------- randomized vs the oracle
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_state_random_ragged(seed):
    n_list = [1, 2, 3, 5, 6, 8, 11, 13]
    progs, _ = _batch(n_list, 100 * seed)
    progs.append(cq.serialize([]))           # empty program
    vals = np.random.default_rng(seed).uniform(0, 2, (len(progs), 3)) \
        .astype(np.float32)
    a = ops.tfq_simulate_state(progs, ["a", "b", "c"], vals)
    b = orc.simulate_state(progs, ["a", "b", "c"], vals)
    assert a.shape == b.shape
    np.testing.assert_allclose(a, b, atol=2e-6)


@pytest.mark.parametrize("seed", [4, 5])
def test_expectation_random_ragged(seed):
    n_list = [2, 3, 4, 7, 9, 12, 14]
    progs, qss = _batch(n_list, 100 * seed)
    sums = _sums(qss, 3, seed)
    progs.append(cq.serialize([]))
    sums.append([cq.pauli_sum([(1.0, [])])] * 3)
    vals = np.random.default_rng(seed).uniform(0, 2, (len(progs), 3)) \
        .astype(np.float32)
    a = ops.tfq_simulate_expectation(progs, ["a", "b", "c"], vals, sums)
    b = orc.simulate_expectation(progs, ["a", "b", "c"], vals, sums)
    np.testing.assert_allclose(a, b, atol=ATOL, rtol=RTOL)
    assert (a[-1] == -2.0).all()


def test_expectation_shared_program_many_rows():
    """PQC-style batch: one program, per-row symbol values, chunked."""
    moments, names, qs = cq.hea_circuit(8, 2)
    prog = cq.serialize(moments)
    B = 37
    vals = np.random.default_rng(7).uniform(0, 2, (B, len(names))) \
        .astype(np.float32)
    obs = cq.hea_observables(qs)
    a = ops.tfq_simulate_expectation([prog] * B, names, vals, [obs] * B)
    b = orc.simulate_expectation([prog] * B, names, vals, [obs] * B)
    np.testing.assert_allclose(a, b, atol=ATOL, rtol=RTOL)
    # small memory budget -> several chunks, same answer
    ctx = ops.get_context()
    ctx.set_memory_budget(5 * (8 << 8) + 100000)
    try:
        c = ops.tfq_simulate_expectation([prog] * B, names, vals, [obs] * B)
    finally:
        ctx.set_memory_budget(0)
    np.testing.assert_array_equal(a, c)


def test_expectation_tile_crossing_qubits():
    """n > tile (12 bits): dense gates on high qubits, controls everywhere."""
    n = 15
    qs = [cq.grid(0, i) for i in range(n)]
    m = cq.random_circuit(qs, 10, 99, controls=True, symbols=("a",))
    sums = [[cq.random_pauli_sum(qs, 6, 5, max_weight=5),
             cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs])]]
    vals = np.array([[0.37]], np.float32)
    a = ops.tfq_simulate_expectation([cq.serialize(m)], ["a"], vals, sums)
    b = orc.simulate_expectation([cq.serialize(m)], ["a"], vals, sums)
    np.testing.assert_allclose(a, b, atol=ATOL, rtol=RTOL)


@pytest.mark.parametrize("seed", [6, 7])
def test_adjoint_random_ragged(seed):
    n_list = [2, 3, 5, 8, 10, 13]
    progs, qss = _batch(n_list, 100 * seed, n_moments=6)
    sums = _sums(qss, 2, seed)
    progs.append(cq.serialize([]))
    sums.append([cq.pauli_sum([(1.0, [])])] * 2)
    rng = np.random.default_rng(seed)
    vals = rng.uniform(0, 2, (len(progs), 3)).astype(np.float32)
    down = rng.normal(size=(len(progs), 2)).astype(np.float32)
    a = ops.tfq_adj_grad(progs, ["a", "b", "c"], vals, sums, down)
    b = orc.adjoint_gradient(progs, ["a", "b", "c"], vals, sums, down)
    # gradient gates carry a 1/(2 eps) = 100x amplification of float32
    # round-off; both sides build identical gradient gates,
    # so the residual is the sweep order only.
    np.testing.assert_allclose(a, b, atol=ATOL, rtol=RTOL)
    assert (a[-1] == 0).all()


def test_adjoint_tfi_and_hea_shared_program():
    for moments, names, qs, obs in (
            (*cq.tfi_chain_circuit(8), None),
            (*cq.hea_circuit(7, 2), "hea")):
        prog = cq.serialize(moments)
        sums = cq.hea_observables(qs) if obs else [cq.tfi_hamiltonian(qs)]
        B = 9
        rng = np.random.default_rng(11)
        vals = rng.uniform(0, 1, (B, len(names))).astype(np.float32)
        down = np.ones((B, len(sums)), np.float32)
        a = ops.tfq_adj_grad([prog] * B, names, vals, [sums] * B, down)
        b = orc.adjoint_gradient([prog] * B, names, vals, [sums] * B, down)
        np.testing.assert_allclose(a, b, atol=ATOL, rtol=RTOL)


def test_samples_bit_exact_with_uniforms_and_padding():
    """Bit-exact sampled bitstrings given identical uniforms AND the identical
    state: the oracle sampler runs on the state the GPU itself exported
    (float32 round-off between two simulators would otherwise move a CDF
    boundary across a uniform about once per 1e3 shots at 12 qubits)."""
    n_list = [1, 3, 6, 9, 12]
    progs, _ = _batch(n_list, 321, controls=False, symbols=())
    progs.append(cq.serialize([]))
    S = 500
    vals = np.zeros((len(progs), 0), np.float32)
    u = np.random.default_rng(5).random((len(progs), S))
    nq = orc.num_qubits(progs)
    states = ops.tfq_simulate_state(progs, [], vals)
    a = ops.tfq_simulate_samples(progs, [], vals, [S], uniforms=u)
    b = orc.samples_from_states(states, nq, S, u)
    assert a.shape == b.shape == (len(progs), S, 12)
    np.testing.assert_array_equal(a, b)
    assert (a[-1] == -2).all()
    # device Philox stream == the oracle's restatement of it
    a = ops.tfq_simulate_samples(progs, [], vals, [S], seed=1234567)
    up = np.stack([orc.philox_uniforms(1234567, i, 0, 0, S)
                   for i in range(len(progs))])
    np.testing.assert_array_equal(a, orc.samples_from_states(states, nq, S, up))
    # end to end against the oracle's own simulation: identical up to the
    # rare boundary crossings caused by float32 state round-off
    c = orc.simulate_samples(progs, [], vals, [S], seed=1234567)
    differing_shots = (a != c).any(axis=2).mean()
    assert differing_shots < 0.01


def test_samples_reference_padding_golden():
    """tfq_simulate_ops_test.py:319-346."""
    progs = [cq.serialize([[cq.X(cq.grid(0, i)) for i in range(n)]])
             for n in (3, 5)]
    out = ops.tfq_simulate_samples(progs, [], np.zeros((2, 0), np.float32), [4])
    assert out.shape == (2, 4, 5)
    assert (out[1] == 1).all()
    assert (out[0, :, :2] == -2).all() and (out[0, :, 2:] == 1).all()
    out = ops.tfq_simulate_samples(progs, [], np.zeros((2, 0), np.float32), [0])
    assert out.shape == (2, 0, 5)


def test_sampled_expectation_exact_with_uniforms():
    n_list = [2, 4, 7, 10]
    progs, qss = _batch(n_list, 555, controls=False, symbols=())
    sums = _sums(qss, 2, 3, max_terms=3)
    vals = np.zeros((len(progs), 0), np.float32)
    ns = np.array([[50, 64], [100, 10], [7, 128], [128, 33]], np.int32)
    u = np.random.default_rng(8).random((len(progs), 2, 3, 128))
    a = ops.tfq_simulate_sampled_expectation(progs, [], vals, sums, ns,
                                             uniforms=u)
    b = orc.simulate_sampled_expectation(progs, [], vals, sums, ns, uniforms=u)
    np.testing.assert_allclose(a, b, atol=1e-6)
    a = ops.tfq_simulate_sampled_expectation(progs, [], vals, sums, ns, seed=99)
    b = orc.simulate_sampled_expectation(progs, [], vals, sums, ns, seed=99)
    np.testing.assert_allclose(a, b, atol=1e-6)


def test_sampled_expectation_converges_to_exact():
    """util_qsim_test.cc:49-116 style: many shots -> analytic value."""
    q0, q1 = cq.grid(0, 0), cq.grid(0, 1)
    circuit = [[cq.X(q0, 0.25)], [cq.CNOT(q0, q1)], [cq.Y(q1, 0.5)]]
    ps = cq.pauli_sum([(0.1234, [(q0, "Z"), (q1, "X")]), (-3.0, [(q0, "X")]),
                       (4.0, [])])
    out = ops.tfq_simulate_sampled_expectation(
        [cq.serialize(circuit)], [], np.zeros((1, 0), np.float32), [[ps]],
        [[200000]], seed=3)
    assert abs(out[0, 0] - 4.1234) < 2e-2


# ----------------------------------------------------------- error strings
def test_error_strings_match_reference():
    q = cq.grid(0, 0)
    prog = cq.serialize([[cq.X(q, "alpha")]])
    z = cq.pauli_sum([(1.0, [(q, "Z")])])
    E = ops.InvalidArgumentError
    with pytest.raises(E, match="Unparseable proto"):
        ops.tfq_simulate_expectation([b"\xff\xfe junk"], ["alpha"], [[0.1]], [[z]])
    with pytest.raises(E, match="Could not find symbol in parameter map"):
        ops.tfq_simulate_expectation([prog], ["beta"], [[0.1]], [[z]])
    with pytest.raises(E, match="qubits not found in circuit"):
        ops.tfq_simulate_expectation(
            [prog], ["alpha"], [[0.1]],
            [[cq.pauli_sum([(1.0, [(cq.grid(5, 5), "Z")])])]])
    with pytest.raises(E, match="do not match"):
        ops.tfq_simulate_expectation([prog, prog], ["alpha"], [[0.1]], [[z], [z]])
    with pytest.raises(E, match="do not match"):
        ops.tfq_simulate_expectation([prog], ["alpha"], [[0.1]], [[z], [z]])
    with pytest.raises(E, match="programs must be rank 1"):
        ops.tfq_simulate_expectation([[prog]], ["alpha"], [[0.1]], [[z]])
    with pytest.raises(E, match="symbol_names must be rank 1"):
        ops.tfq_simulate_expectation([prog], [["alpha"]], [[0.1]], [[z]])
    with pytest.raises(E, match="symbol_values must be rank 2"):
        ops.tfq_simulate_expectation([prog], ["alpha"], [0.1], [[z]])
    with pytest.raises(E, match="pauli_sums must be rank 2"):
        ops.tfq_simulate_expectation([prog], ["alpha"], [[0.1]], [z])
    with pytest.raises(E, match="gradients and circuits do not match"):
        ops.tfq_adj_grad([prog], ["alpha"], [[0.1]], [[z]],
                         np.ones((2, 1), np.float32))
    with pytest.raises(E, match="gradients and pauli sum dimension do not match"):
        ops.tfq_adj_grad([prog], ["alpha"], [[0.1]], [[z]],
                         np.ones((1, 2), np.float32))
    with pytest.raises(E, match="greater than 0"):
        ops.tfq_simulate_sampled_expectation([prog], ["alpha"], [[0.1]], [[z]],
                                             [[0]])
    with pytest.raises(E, match="num_samples must be rank 2"):
        ops.tfq_simulate_sampled_expectation([prog], ["alpha"], [[0.1]], [[z]],
                                             [1])
    bad = cq.to_program([[cq.X(q, 0.5)]])
    bad.circuit.moments[0].operations[0].gate.id = "ADP"
    with pytest.raises(E, match="cirq.Channel"):
        ops.tfq_simulate_expectation([bad.SerializeToString()], [],
                                     np.zeros((1, 0), np.float32), [[z]])


# ------------------------------------------- full-size, property-based checks
def test_full_size_20q_properties():
    """BASELINE configs[1] size (20 qubits): norm preservation, <Z>-sum
    bounds, adjoint gradient vs a finite difference of the GPU expectation."""
    moments, names, qs = cq.hea_circuit(20, 4)
    prog = cq.serialize(moments)
    obs = cq.hea_observables(qs)
    B = 4
    rng = np.random.default_rng(20)
    vals = rng.uniform(0, 2, (B, len(names))).astype(np.float32)
    e = ops.tfq_simulate_expectation([prog] * B, names, vals, [obs] * B)
    assert np.isfinite(e).all() and (np.abs(e[:, 0]) <= 20 + 1e-3).all()
    # identity-only observable = norm of the state
    one = [[cq.pauli_sum([(1.0, [(qs[0], "Z"), (qs[0], "Z")])])]] * B
    # ZZ on one qubit is not producible by the serializer; use <I> instead
    one = [[cq.pauli_sum([(1.0, [])])]] * B
    assert np.allclose(ops.tfq_simulate_expectation([prog] * B, names, vals, one), 1.0)
    # oracle agreement on one row at full size (C executor: seconds)
    b = orc.simulate_expectation([prog], names, vals[:1], [obs])
    np.testing.assert_allclose(e[:1], b, atol=ATOL, rtol=RTOL)
    g = ops.tfq_adj_grad([prog] * 2, names, vals[:2], [obs] * 2,
                         np.ones((2, 4), np.float32))
    gb = orc.adjoint_gradient([prog], names, vals[:1], [obs],
                              np.ones((1, 4), np.float32))
    np.testing.assert_allclose(g[:1], gb, atol=ATOL, rtol=RTOL)
    # central difference of the forward op along 3 symbols
    h = 1e-2
    for col in (0, 57, 159):
        vp, vm = vals[:1].copy(), vals[:1].copy()
        vp[0, col] += h
        vm[0, col] -= h
        ep = ops.tfq_simulate_expectation([prog], names, vp, [obs]).sum()
        em = ops.tfq_simulate_expectation([prog], names, vm, [obs]).sum()
        assert abs((ep - em) / (2 * h) - g[0, col]) < 2e-2


# ------------------------------------------- sharded single state
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_state_matches_unsharded(world):
    """One state split over `world` virtual ranks on this GPU (the exchange is
    emulated by chunk copies; the NCCL driver shares every other line of code):
    expectations equal the unsharded op and the oracle."""
    from quantum_b200 import sharded
    n = 13
    qs = [cq.grid(0, i) for i in range(n)]
    m = cq.random_circuit(qs, 12, 4242, controls=True, symbols=("a", "b"))
    prog = cq.serialize(m)
    sums = [cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs]),
            cq.pauli_sum([(0.7, [(q, "X")]) for q in qs] +
                         [(0.5, [(qs[i], "X"), (qs[i + 1], "Y")])
                          for i in range(n - 1)] + [(0.25, [])]),
            cq.random_pauli_sum(qs, 6, 9, max_weight=4)]
    vals = np.array([[0.37, 1.21]], np.float32)
    stats = {}
    a = sharded.emulated_sharded_expectation(prog, ["a", "b"], vals[0], sums,
                                             world, stats=stats)
    b = ops.tfq_simulate_expectation([prog], ["a", "b"], vals, [sums])[0]
    c = orc.simulate_expectation([prog], ["a", "b"], vals, [sums])[0]
    assert stats["exchanges"] >= 1        # the path under test really swaps
    np.testing.assert_allclose(a, b, atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(a, c, atol=ATOL, rtol=RTOL)


def test_sharded_state_c5_style_circuit():
    from quantum_b200 import sharded
    m, qs = cq.supremacy_style_circuit(4, 4, 20, 16, use_line=True)
    prog = cq.serialize(m)
    sums = [cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs])]
    a = sharded.emulated_sharded_expectation(prog, [], np.zeros(0, np.float32),
                                             sums, 8)
    b = ops.tfq_simulate_expectation([prog], [], np.zeros((1, 0), np.float32),
                                     [sums])[0]
    np.testing.assert_allclose(a, b, atol=ATOL, rtol=RTOL)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_state_peer_memory_exchange(world):
    """The exchange inside the library (tfqb_sharded_export / _connect /
    _enqueue / _result): every virtual rank has its own stream, the qubit
    swaps are peer-memory pulls ordered by epoch flags, the partial sums are
    added in rank order on every rank.  Two evaluations per job: the flags
    keep counting, the buffers swap roles."""
    from quantum_b200 import sharded
    n = 13
    qs = [cq.grid(0, i) for i in range(n)]
    m = cq.random_circuit(qs, 12, 4242, controls=True, symbols=("a", "b"))
    prog = cq.serialize(m)
    sums = [cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs]),
            cq.pauli_sum([(0.7, [(q, "X")]) for q in qs] +
                         [(0.5, [(qs[i], "X"), (qs[i + 1], "Y")])
                          for i in range(n - 1)] + [(0.25, [])])]
    vals = np.array([[0.37, 1.21]], np.float32)
    stats = {}
    outs = sharded.emulated_peer_sharded_expectation(prog, ["a", "b"], vals[0], sums,
                                                     world, repeats=2, stats=stats)
    assert len(outs) == 2 * world and stats["exchanges"] >= 1
    b = ops.tfq_simulate_expectation([prog], ["a", "b"], vals, [sums])[0]
    for o in outs:
        np.testing.assert_array_equal(o, outs[0])     # identical bits on every rank
    np.testing.assert_allclose(outs[0], b, atol=ATOL, rtol=RTOL)
    # the same stages with the host-driven exchange agree
    a = sharded.emulated_sharded_expectation(prog, ["a", "b"], vals[0], sums, world)
    np.testing.assert_allclose(outs[0], a, atol=1e-6)


def test_sharded_state_peer_memory_standalone_pull():
    """TFQB_FUSED_EXCHANGE=0: the qubit swap as its own pull kernel instead of
    the load phase of the next gate pass (the default, covered above)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import json, sys
import numpy as np
sys.path.insert(0, %r)
from quantum_b200 import circuits as cq, ops, sharded
qs = [cq.grid(0, i) for i in range(13)]
prog = cq.serialize(cq.random_circuit(qs, 12, 4242, controls=True, symbols=("a", "b")))
sums = [cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs]),
        cq.pauli_sum([(0.7, [(q, "X")]) for q in qs])]
vals = np.array([[0.37, 1.21]], np.float32)
st = {}
outs = sharded.emulated_peer_sharded_expectation(prog, ["a", "b"], vals[0], sums, 4,
                                                 repeats=2, stats=st)
ref = ops.tfq_simulate_expectation([prog], ["a", "b"], vals, [sums])[0]
print(json.dumps({"err": float(np.abs(outs[0] - ref).max()),
                  "same": bool(all((o == outs[0]).all() for o in outs)),
                  "exchanges": st["exchanges"], "fused": st["fused_exchanges"]}))
""" % root
    for fused in ("0", "1"):
        env = dict(os.environ, TFQB_FUSED_EXCHANGE=fused)
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                             timeout=600, env=env)
        assert res.returncode == 0, res.stderr[-2000:]
        out = json.loads(res.stdout.strip().splitlines()[-1])
        assert out["same"] and out["err"] < 2e-5 and out["exchanges"] >= 1
        assert (out["fused"] > 0) == (fused == "1")


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_state_sampling(world):
    """tfqb_sharded_sample: shard norms through the flag blocks, every shot
    owned by exactly one rank, and the bitstrings are what the same sampler
    draws from the (unsharded) state laid out in the shards' physical order
    with the same uniforms; marginals match the exact probabilities."""
    from quantum_b200 import sharded
    n, S = 13, 3000
    qs = [cq.grid(0, i) for i in range(n)]
    prog = cq.serialize(cq.random_circuit(qs, 12, 777, controls=False, symbols=()))
    vals = np.zeros((1, 0), np.float32)
    u = np.random.default_rng(2).random(S)
    got, count = sharded.emulated_peer_sharded_samples(prog, [], vals[0], S, n, world,
                                                       uniforms=u)
    assert (count == 1).all()
    assert set(np.unique(got)) <= {0, 1}
    psi = ops.tfq_simulate_state([prog], [], vals)[0]
    prob = np.abs(psi.astype(np.complex128)) ** 2
    # exact marginals P(bit q = 1), output column n-1-q
    for q in range(n):
        p1 = prob[(np.arange(2 ** n) >> q) & 1 == 1].sum()
        f = got[:, n - 1 - q].mean()
        assert abs(f - p1) < 4.5 * np.sqrt(max(p1 * (1 - p1), 1e-4) / S) + 1e-3, (q, f, p1)
    # the same draw from the oracle's sampler over the physical amplitude order
    plan = ops.host_describe_sharded(prog, [], [], world)
    phys = plan["final_phys"]
    idx = np.arange(2 ** n)
    pidx = np.zeros_like(idx)
    for b in range(n):
        pidx |= ((idx >> b) & 1) << phys[b]
    state_phys = np.zeros(2 ** n, np.complex64)
    state_phys[pidx] = psi
    nl = plan["n_local"]
    norms = np.array([(np.abs(state_phys[r << nl:(r + 1) << nl].astype(np.complex128)) ** 2).sum()
                      for r in range(world)])
    cum = np.concatenate([[0.0], np.cumsum(norms)])
    us = np.sort(u)
    want = np.zeros((S, n), np.int8)
    for s_, x in enumerate(us * cum[-1]):
        r = min(int(np.searchsorted(cum[1:], x, side="right")), world - 1)
        v = min(max((x - cum[r]) / norms[r], 0.0), np.nextafter(1.0, 0.0))
        local = int(orc.sample_tree(state_phys[r << nl:(r + 1) << nl], np.array([v]))[0])
        p = (r << nl) | local
        for b in range(n):
            want[s_, n - 1 - b] = (p >> phys[b]) & 1
    # identical except where float32 differences between the sharded and the
    # unsharded simulation move a boundary across a uniform
    assert (got != want).any(axis=1).mean() < 0.01


def test_sharded_state_peer_memory_two_gpus():
    """Real peer memory: 2 processes, CUDA IPC over NVLink."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(root, "scripts", "bench_sharded.py"),
           "--qubits", "20", "--reps", "2", "--check", "--xterms", "--exchange", "peer"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["exchanges"] >= 1
    assert abs(line["expectation"] - line["unsharded_check"]) < 1e-4


def test_sharded_state_over_nccl_two_gpus():
    """Real exchange: 2 ranks, torch.distributed NCCL all_to_all_single."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "scripts", "bench_sharded.py"),
           "--qubits", "20", "--reps", "1", "--check", "--xterms", "--exchange", "nccl"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["exchanges"] >= 1
    assert abs(line["expectation"] - line["unsharded_check"]) < 1e-4


# ------------------------------------------------ chunked execution (streaming)
def test_all_ops_identical_when_streamed_in_chunks():
    """A small memory budget forces several chunks per group for every op;
    results must not change (BASELINE config 4 streams 16384 x 22q rows)."""
    moments, names, qs = cq.tfi_chain_circuit(9)
    prog = cq.serialize(moments)
    ham = [cq.tfi_hamiltonian(qs)]
    B = 23
    rng = np.random.default_rng(3)
    vals = rng.uniform(0, 1, (B, len(names))).astype(np.float32)
    down = rng.normal(size=(B, 1)).astype(np.float32)
    u = rng.random((B, 32))
    ns = np.full((B, 1), 40, np.int32)

    def run_all():
        return (ops.tfq_simulate_expectation([prog] * B, names, vals, [ham] * B),
                ops.tfq_adj_grad([prog] * B, names, vals, [ham] * B, down),
                ops.tfq_simulate_state([prog] * B, names, vals),
                ops.tfq_simulate_samples([prog] * B, names, vals, [32], uniforms=u),
                ops.tfq_simulate_sampled_expectation([prog] * B, names, vals,
                                                     [ham] * B, ns, seed=5))
    full = run_all()
    ctx = ops.get_context()
    ctx.set_memory_budget(4 * 3 * (8 << 9) + 200000)   # ~4 rows of 3 buffers
    try:
        chunked = run_all()
    finally:
        ctx.set_memory_budget(0)
    for a, b in zip(full, chunked):
        np.testing.assert_array_equal(a, b)
    ref = orc.adjoint_gradient([prog] * B, names, vals, [ham] * B, down)
    np.testing.assert_allclose(full[1], ref, atol=ATOL, rtol=RTOL)


def test_product_state_and_sign_op_paths():
    """Circuits that are only 1-qubit gates (pass 0 synthesises the product
    state and applies nothing), literal CZ/Z/ZZ gates (sign ops), and a
    leading 2-qubit gate (no product init on those qubits)."""
    qs = [cq.grid(0, i) for i in range(7)]
    c1 = [[cq.X(q, 0.3 + 0.1 * i) for i, q in enumerate(qs)],
          [cq.Z(q, 0.7) for q in qs]]
    c2 = [[cq.H(q) for q in qs], [cq.CZ(qs[i], qs[i + 1]) for i in range(0, 6, 2)],
          [cq.Z(qs[1]), cq.ZZ(qs[2], qs[3]), cq.Y(qs[5], 0.25)],
          [cq.CZ(qs[i], qs[i + 1]) for i in range(1, 6, 2)], [cq.H(q) for q in qs]]
    c3 = [[cq.ISWAP(qs[0], qs[1], 0.4), cq.H(qs[2])], [cq.X(q, "a") for q in qs],
          [cq.CNOT(qs[6], qs[0])]]
    progs = [cq.serialize(c) for c in (c1, c2, c3)]
    sums = [[cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs] + [(0.5, [(qs[0], "X"), (qs[6], "Y")])])]] * 3
    vals = np.array([[0.4], [0.4], [0.4]], np.float32)
    a = ops.tfq_simulate_state(progs, ["a"], vals)
    b = orc.simulate_state(progs, ["a"], vals)
    np.testing.assert_allclose(a, b, atol=2e-6)
    e = ops.tfq_simulate_expectation(progs, ["a"], vals, sums)
    f = orc.simulate_expectation(progs, ["a"], vals, sums)
    np.testing.assert_allclose(e, f, atol=ATOL, rtol=RTOL)
    g = ops.tfq_adj_grad(progs, ["a"], vals, sums, np.ones((3, 1), np.float32))
    h = orc.adjoint_gradient(progs, ["a"], vals, sums, np.ones((3, 1), np.float32))
    np.testing.assert_allclose(g, h, atol=ATOL, rtol=RTOL)


def test_deterministic_switch_is_bit_reproducible():
    """TFQB_DETERMINISTIC=1: one CTA per row sums the gradient slots and the
    per-term partials in a fixed order, so two runs agree bit for bit (the
    default path adds per-CTA fp64 partials with atomics: last-bit noise), and
    the values are those of the default path to rounding."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import json, sys
import numpy as np
sys.path.insert(0, %r)
from quantum_b200 import circuits as cq, ops
mo, names, qs = cq.hea_circuit(16, 3)
prog = cq.serialize(mo)
obs = cq.hea_observables(qs)
B = 24
v = np.random.default_rng(3).uniform(0, 2, (B, len(names))).astype(np.float32)
runs = []
for _ in range(3):
    e = ops.tfq_simulate_expectation([prog] * B, names, v, [obs] * B)
    g = ops.tfq_adj_grad([prog] * B, names, v, [obs] * B, np.ones((B, 4), np.float32))
    runs.append((e, g))
same = all(np.array_equal(runs[0][0], r[0]) and np.array_equal(runs[0][1], r[1]) for r in runs[1:])
np.save(sys.argv[1], np.concatenate([runs[-1][0].ravel(), runs[-1][1].ravel()]))
print(json.dumps({"same": bool(same), "jit": ops.get_context().profile_read()["jit_pass_launches"]}))
''' % root
    outs = {}
    for tag, extra in (("det", {"TFQB_DETERMINISTIC": "1"}), ("default", {})):
        env = dict(os.environ, TFQB_JIT_MIN_AMPS="0", **extra)
        path = os.path.join("/tmp", "tfqb_det_%s_%d.npy" % (tag, os.getpid()))
        res = subprocess.run([sys.executable, "-c", code, path], capture_output=True,
                             text=True, timeout=900, env=env)
        assert res.returncode == 0, res.stderr[-3000:]
        outs[tag] = (json.loads(res.stdout.strip().splitlines()[-1]), np.load(path))
        os.remove(path)
    assert outs["det"][0]["same"] and outs["det"][0]["jit"] > 0
    np.testing.assert_allclose(outs["det"][1], outs["default"][1], atol=2e-5, rtol=1e-4)


def test_committed_golden_fixtures():
    """The CUDA path against the committed fixtures (tests/golden/*.npz)."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(here, "five_ops_ragged.npz"), allow_pickle=True)
    progs = list(z["programs"])
    names = list(z["symbol_names"])
    sums = [list(r) for r in z["pauli_sums"]]
    vals = z["symbol_values"]
    np.testing.assert_allclose(ops.tfq_simulate_state(progs, names, vals),
                               z["state"], atol=2e-6)
    np.testing.assert_allclose(
        ops.tfq_simulate_expectation(progs, names, vals, sums), z["expectation"],
        atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(
        ops.tfq_adj_grad(progs, names, vals, sums, z["downstream"]), z["gradient"],
        atol=ATOL, rtol=RTOL)
    got = ops.tfq_simulate_samples(progs, names, vals, [64], uniforms=z["uniforms"])
    assert (got != z["samples"]).any(axis=2).mean() < 0.02   # float32 state round-off
    se = ops.tfq_simulate_sampled_expectation(progs, names, vals, sums,
                                              z["num_samples"],
                                              uniforms=z["uniforms_exp"])
    # identical except where float32 state round-off moves a CDF boundary
    # across a uniform (about 1 shot in 1e3): at most 3 of the 14 entries
    assert (np.abs(se - z["sampled_expectation"]) > 1e-6).sum() <= 3
    h = np.load(os.path.join(here, "hea12_expectation_adjoint.npz"), allow_pickle=True)
    prog, hn = h["program"][0], list(h["symbol_names"])
    obs = list(h["pauli_sums"])
    v = h["symbol_values"]
    np.testing.assert_allclose(
        ops.tfq_simulate_expectation([prog] * 5, hn, v, [obs] * 5), h["expectation"],
        atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(
        ops.tfq_adj_grad([prog] * 5, hn, v, [obs] * 5, np.ones((5, 4), np.float32)),
        h["gradient"], atol=ATOL, rtol=RTOL)


# ------------------------------------------------ N1: inner product (next row)
def test_inner_product_matches_oracle_and_error_strings():
    """TfqInnerProduct (math_ops/tfq_inner_product.cc:45-292)."""
    n_list = [2, 4, 7, 11, 13]
    progs, others = [], []
    for k, n in enumerate(n_list):
        qs = [cq.grid(0, i) for i in range(n)]
        progs.append(cq.serialize(cq.random_circuit(qs, 8, 900 + k, controls=True,
                                                    symbols=("a", "b"))))
        row = []
        for j in range(3):
            m = cq.random_circuit(qs, 5, 50 * k + j, controls=(j == 1))
            m.append([cq.H(q) for q in qs])          # touches every qubit
            row.append(cq.serialize(m))
        others.append(row)
    progs.append(cq.serialize([]))
    others.append([cq.serialize([])] * 3)
    vals = np.random.default_rng(1).uniform(0, 2, (len(progs), 2)).astype(np.float32)
    a = ops.tfq_inner_product(progs, ["a", "b"], vals, others)
    b = orc.inner_product(progs, ["a", "b"], vals, others)
    assert a.shape == b.shape == (len(progs), 3)
    np.testing.assert_allclose(a, b, atol=ATOL, rtol=RTOL)
    assert (a[-1] == 1).all()
    # <psi|psi> = 1 when the paired circuit is the (symbol-resolved) circuit
    qs = [cq.grid(0, i) for i in range(9)]
    m = cq.random_circuit(qs, 6, 5)
    m.append([cq.H(q) for q in qs])
    same = cq.serialize(m)
    one = ops.tfq_inner_product([same] * 4, [], np.zeros((4, 0), np.float32),
                                [[same]] * 4)
    np.testing.assert_allclose(one, np.ones((4, 1)), atol=2e-6)
    E = ops.InvalidArgumentError
    q0, q1 = cq.grid(0, 0), cq.grid(0, 1)
    ref = cq.serialize([[cq.X(q0), cq.X(q1)]])
    with pytest.raises(E, match="qubits not found in reference circuit"):
        ops.tfq_inner_product([ref], [], np.zeros((1, 0), np.float32),
                              [[cq.serialize([[cq.X(q0), cq.X(cq.grid(5, 5))]])]])
    with pytest.raises(E, match="qubits not found in paired circuit"):
        ops.tfq_inner_product([ref], [], np.zeros((1, 0), np.float32),
                              [[cq.serialize([[cq.X(q0)]])]])
    with pytest.raises(E, match="Found symbols in other_programs"):
        ops.tfq_inner_product([ref], [], np.zeros((1, 0), np.float32),
                              [[cq.serialize([[cq.X(q0, "s"), cq.X(q1)]])]])
    with pytest.raises(E, match="batch dimension do not match"):
        ops.tfq_inner_product([ref], [], np.zeros((1, 0), np.float32), [[ref], [ref]])
    with pytest.raises(E, match="other_programs must be rank 2"):
        ops.tfq_inner_product([ref], [], np.zeros((1, 0), np.float32), [ref])


def test_inner_product_grad_matches_oracle_and_error_strings():
    """TfqInnerProductGrad (math_ops/tfq_inner_product_grad.cc:46-501): the
    complex gradient of <psi(theta)|phi> weighted by the downstream gradients,
    from two real reverse sweeps (lam and i*lam)."""
    n_list = [2, 4, 7, 11, 13]
    progs, others = [], []
    for k, n in enumerate(n_list):
        qs = [cq.grid(0, i) for i in range(n)]
        progs.append(cq.serialize(cq.random_circuit(qs, 8, 700 + k, controls=True,
                                                    symbols=("a", "b", "c"))))
        row = []
        for j in range(2):
            m = cq.random_circuit(qs, 5, 30 * k + j, controls=(j == 1))
            m.append([cq.H(q) for q in qs])          # touches every qubit
            row.append(cq.serialize(m))
        others.append(row)
    progs.append(cq.serialize([]))
    others.append([cq.serialize([])] * 2)
    rng = np.random.default_rng(3)
    vals = rng.uniform(0, 2, (len(progs), 3)).astype(np.float32)
    down = rng.uniform(-1, 1, (len(progs), 2)).astype(np.float32)
    a = ops.tfq_inner_product_grad(progs, ["a", "b", "c"], vals, others, down)
    b = orc.inner_product_grad(progs, ["a", "b", "c"], vals, others, down)
    assert a.shape == b.shape == (len(progs), 3) and a.dtype == np.complex64
    # finite-difference gradient gates amplify float32 round-off by 100
    np.testing.assert_allclose(a, b, atol=ATOL, rtol=RTOL)
    assert (a[-1] == 0).all()
    # the gradient of <psi(theta)|phi> itself: central differences of the
    # forward op through the same library
    q = [cq.grid(0, i) for i in range(5)]
    circ = [[cq.H(x) for x in q], [cq.X(q[0], "a"), cq.Y(q[1], "a"), cq.ZZ(q[2], q[3], "b")],
            [cq.CNOT(q[0], q[4]), cq.FSim(q[1], q[2], 0.3, "b")]]
    prog = cq.serialize(circ)
    oth = cq.serialize(cq.random_circuit(q, 6, 11) + [[cq.H(x) for x in q]])
    v = np.array([[0.37, 1.21]], np.float32)
    g = ops.tfq_inner_product_grad([prog], ["a", "b"], v, [[oth]], np.ones((1, 1), np.float32))
    for col in range(2):
        dv = np.zeros_like(v)
        dv[0, col] = 1e-2
        fd = (ops.tfq_inner_product([prog], ["a", "b"], v + dv, [[oth]]) -
              ops.tfq_inner_product([prog], ["a", "b"], v - dv, [[oth]]))[0, 0] / 2e-2
        assert abs(g[0, col] - fd) < 2e-3
    E = ops.InvalidArgumentError
    q0, q1 = cq.grid(0, 0), cq.grid(0, 1)
    ref = cq.serialize([[cq.X(q0, "a"), cq.X(q1)]])
    oth = cq.serialize([[cq.X(q0), cq.X(q1)]])
    one = np.ones((1, 1), np.float32)
    with pytest.raises(E, match="number of symbols must be a positive integer"):
        ops.tfq_inner_product_grad([oth], [], np.zeros((1, 0), np.float32), [[oth]], one)
    with pytest.raises(E, match="gradients and circuits do not match"):
        ops.tfq_inner_product_grad([ref], ["a"], np.zeros((1, 1), np.float32), [[oth]],
                                   np.ones((2, 1), np.float32))
    with pytest.raises(E, match="gradients and other_programs do not match"):
        ops.tfq_inner_product_grad([ref], ["a"], np.zeros((1, 1), np.float32), [[oth]],
                                   np.ones((1, 2), np.float32))
    with pytest.raises(E, match="Found symbols in other_programs"):
        ops.tfq_inner_product_grad([ref], ["a"], np.zeros((1, 1), np.float32), [[ref]], one)
    with pytest.raises(E, match="other_programs must be rank 2"):
        ops.tfq_inner_product_grad([ref], ["a"], np.zeros((1, 1), np.float32), [oth], one)


def test_specialised_pass_kernels_parity():
    """TFQB_JIT_MIN_AMPS=0 compiles the run-time specialised pass kernels
    (csrc/jit.cc: NVRTC, same device primitives) on the first call.  States,
    expectations and gradients against the oracle, forward and adjoint, on
    the HEA / TFI workloads and on random circuits (whose controlled passes
    stay on the interpreted kernel), and the profile counters prove that the
    specialised kernels ran."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import json, sys
import numpy as np
sys.path.insert(0, %r)
from oracle import tfq_oracle as orc
from quantum_b200 import circuits as cq, ops
out = {}
ctx = ops.get_context()
mo, names, q2 = cq.hea_circuit(14, 3)
p2 = cq.serialize(mo)
v2 = np.random.default_rng(1).uniform(0, 2, (3, len(names))).astype(np.float32)
obs = cq.hea_observables(q2)
a = ops.tfq_simulate_state([p2] * 3, names, v2)
b = orc.simulate_state([p2] * 3, names, v2)
out["hea_state_err"] = float(np.abs(a - b).max())
e2 = ops.tfq_simulate_expectation([p2] * 3, names, v2, [obs] * 3)
f2 = orc.simulate_expectation([p2] * 3, names, v2, [obs] * 3)
out["hea_exp_err"] = float(np.abs(e2 - f2).max())
g2 = ops.tfq_adj_grad([p2] * 3, names, v2, [obs] * 3, np.ones((3, 4), np.float32))
h2 = orc.adjoint_gradient([p2] * 3, names, v2, [obs] * 3, np.ones((3, 4), np.float32))
def ratio(a, b):   # worst |a - b| / (1e-5 + 1e-4 |b|): <= 1 is north_star's tolerance
    return float((np.abs(a - b) / (1e-5 + 1e-4 * np.abs(b))).max())
out["hea_grad_err"] = float(np.abs(g2 - h2).max())
out["hea_grad_ratio"] = ratio(g2, h2)
out["hea_grad_scale"] = float(np.abs(h2).max())
mo, names, q3 = cq.tfi_chain_circuit(13, 2)
p3 = cq.serialize(mo)
v3 = np.random.default_rng(2).uniform(0, 1, (2, len(names))).astype(np.float32)
ob3 = [cq.tfi_hamiltonian(q3)]
e3 = ops.tfq_simulate_expectation([p3] * 2, names, v3, [ob3] * 2)
f3 = orc.simulate_expectation([p3] * 2, names, v3, [ob3] * 2)
out["tfi_exp_err"] = float(np.abs(e3 - f3).max())
g3 = ops.tfq_adj_grad([p3] * 2, names, v3, [ob3] * 2, np.ones((2, 1), np.float32))
h3 = orc.adjoint_gradient([p3] * 2, names, v3, [ob3] * 2, np.ones((2, 1), np.float32))
out["tfi_grad_err"] = float(np.abs(g3 - h3).max())
out["tfi_grad_ratio"] = ratio(g3, h3)
for seed, controls in ((11, False), (12, True)):
    qs = [cq.grid(0, i) for i in range(13)]
    m = cq.random_circuit(qs, 10, seed, controls=controls, symbols=("a", "b"))
    prog = cq.serialize(m)
    vals = np.array([[0.3, 1.1], [0.9, 0.2]], np.float32)
    sums = [[cq.random_pauli_sum(qs, 6, 3, max_weight=4),
             cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs])]] * 2
    a = ops.tfq_simulate_state([prog] * 2, ["a", "b"], vals)
    b = orc.simulate_state([prog] * 2, ["a", "b"], vals)
    out["rand%%d_state_err" %% seed] = float(np.abs(a - b).max())
    g = ops.tfq_adj_grad([prog] * 2, ["a", "b"], vals, sums, np.ones((2, 2), np.float32))
    h = orc.adjoint_gradient([prog] * 2, ["a", "b"], vals, sums, np.ones((2, 2), np.float32))
    out["rand%%d_grad_err" %% seed] = float(np.abs(g - h).max())
    out["rand%%d_grad_ratio" %% seed] = ratio(g, h)
# every structured 1-qubit form of the specialised kernels (pass_device.cuh
# g1_*_lift / adj1_*_lift): R D (Z then Y), R alone (Y), D R (Y then Z), X^t,
# H (a reflection: matrix form), exponents beyond +-1 (rotations past pi/2:
# the dropped global sign), CZ layers in between
q4 = [cq.grid(0, i) for i in range(13)]
m4, n4 = [[cq.H(q) for q in q4]], []
def lay(gate, tag):
    ms = []
    for i, q in enumerate(q4):
        n4.append("%%s%%d" %% (tag, i))
        ms.append(gate(q, n4[-1]))
    return ms
for rep in range(2):
    m4 += [lay(cq.Z, "zb%%d_" %% rep), lay(cq.Y, "ya%%d_" %% rep),
           [cq.CZ(q4[i], q4[i + 1]) for i in range(0, 12, 2)],
           lay(cq.Y, "yc%%d_" %% rep), [cq.CZ(q4[i], q4[i + 1]) for i in range(1, 12, 2)],
           lay(cq.X, "xd%%d_" %% rep), [cq.H(q) for q in q4[::3]],
           lay(cq.Y, "ye%%d_" %% rep), lay(cq.Z, "zf%%d_" %% rep),
           [cq.CZ(q4[i], q4[i + 1]) for i in range(0, 12, 2)]]
p4 = cq.serialize(m4)
v4 = np.random.default_rng(4).uniform(-2.5, 2.5, (3, len(n4))).astype(np.float32)
ob4 = cq.hea_observables(q4)
e4 = ops.tfq_simulate_expectation([p4] * 3, n4, v4, [ob4] * 3)
f4 = orc.simulate_expectation([p4] * 3, n4, v4, [ob4] * 3)
out["forms_exp_ratio"] = ratio(e4, f4)
g4 = ops.tfq_adj_grad([p4] * 3, n4, v4, [ob4] * 3, np.ones((3, 4), np.float32))
h4 = orc.adjoint_gradient([p4] * 3, n4, v4, [ob4] * 3, np.ones((3, 4), np.float32))
out["forms_grad_ratio"] = ratio(g4, h4)
out["forms_grad_err"] = float(np.abs(g4 - h4).max())
src = ops.host_jit_source(p4, n4, pass_index=0, phase_free=True) + \
      ops.host_jit_source(p4, n4, adjoint=True, pass_index=0)
out["forms_seen"] = [k for k in ("g1_colreal_lift<", "g1_rowreal_lift<", "g1_real_lift<",
                                 "g1_ximag_lift<", "adj1_real_lift<", "adj1_ximag_lift<")
                     if k in src]
out["profile"] = ctx.profile_read()
print(json.dumps(out))
''' % root
    env = dict(os.environ, TFQB_JIT_MIN_AMPS="0", TFQB_JIT_VERBOSE="1")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True,
                         text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    prof = out["profile"]
    assert prof["jit_kernels"] > 0 and prof["jit_pass_launches"] > 0, res.stderr[-2000:]
    assert out["hea_state_err"] < 2e-6 and out["rand11_state_err"] < 2e-6
    assert out["rand12_state_err"] < 2e-6
    assert out["hea_exp_err"] < ATOL + RTOL and out["tfi_exp_err"] < ATOL + 20 * RTOL
    print(json.dumps({k: v for k, v in out.items() if k != "profile"}))
    assert out["hea_grad_ratio"] <= 1.0 and out["tfi_grad_ratio"] <= 1.0
    assert out["rand11_grad_ratio"] <= 1.0 and out["rand12_grad_ratio"] <= 1.0
    assert out["forms_exp_ratio"] <= 1.0 and out["forms_grad_ratio"] <= 1.0
    assert len(out["forms_seen"]) >= 4, out["forms_seen"]
