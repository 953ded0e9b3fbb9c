"""One context over several GPUs (tfqb_create_multi): every op call fans its
rows over the devices (the reference spreads one Compute over all host cores,
tfq_simulate_expectation_op.cc:245-248) and must return exactly what the
single-device call returns — including the sampling ops, whose Philox streams
are keyed by global row.

On a one-GPU box the same ordinal is listed three times
(TFQB_MULTI_ALLOW_DUPLICATES=1): same fan-out code, same slices, one device.
With two or more GPUs the devices are distinct.
"""
import os

import numpy as np
import pytest

from quantum_b200 import circuits as cq
from quantum_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def multi():
    import torch
    n = torch.cuda.device_count()
    if n >= 2:
        devs = list(range(min(n, 4)))
    else:
        os.environ["TFQB_MULTI_ALLOW_DUPLICATES"] = "1"
        devs = [0, 0, 0]
    ctx = ops.Context(devs)
    assert ctx.device_count() == len(devs)
    ops._contexts[tuple(devs)] = ctx
    yield tuple(devs)
    ops._contexts.pop(tuple(devs), None)
    ctx.close()


def _ragged(seed, n_list):
    progs, qss = [], []
    for k, n in enumerate(n_list):
        qs = [cq.grid(0, i) for i in range(n)]
        progs.append(cq.serialize(cq.random_circuit(qs, 6, seed + k, symbols=("a", "b"))))
        qss.append(qs)
    return progs, qss


def test_multi_matches_single_all_ops(multi):
    n_list = [3, 9, 5, 12, 7, 4, 11, 6, 2, 8]      # widest row in the middle block
    progs, qss = _ragged(77, n_list)
    progs.insert(4, cq.serialize([]))               # an empty program
    qss.insert(4, qss[0])
    B = len(progs)
    names = ["a", "b"]
    vals = np.random.default_rng(1).uniform(0, 2, (B, 2)).astype(np.float32)
    sums = [[cq.random_pauli_sum(qs, 4, 5 + i, max_weight=3),
             cq.pauli_sum([(0.5, [(qs[0], "Z")]), (0.25, [])])] for i, qs in enumerate(qss)]
    down = np.random.default_rng(2).normal(size=(B, 2)).astype(np.float32)
    one = dict(device=0)
    many = dict(device=multi)
    np.testing.assert_array_equal(
        ops.tfq_simulate_expectation(progs, names, vals, sums, **many),
        ops.tfq_simulate_expectation(progs, names, vals, sums, **one))
    np.testing.assert_array_equal(
        ops.tfq_simulate_state(progs, names, vals, **many),
        ops.tfq_simulate_state(progs, names, vals, **one))
    g1 = ops.tfq_adj_grad(progs, names, vals, sums, down, **one)
    gm = ops.tfq_adj_grad(progs, names, vals, sums, down, **many)
    # cross-tile fp64 atomics: run-to-run differences far below float32
    np.testing.assert_allclose(gm, g1, atol=1e-6, rtol=1e-6)
    s1 = ops.tfq_simulate_samples(progs, names, vals, [40], seed=123, **one)
    sm = ops.tfq_simulate_samples(progs, names, vals, [40], seed=123, **many)
    assert sm.shape == (B, 40, 12)
    np.testing.assert_array_equal(sm, s1)
    ns = np.full((B, 2), 64, np.int32)
    np.testing.assert_array_equal(
        ops.tfq_simulate_sampled_expectation(progs, names, vals, sums, ns, seed=9, **many),
        ops.tfq_simulate_sampled_expectation(progs, names, vals, sums, ns, seed=9, **one))


def test_multi_device_jobs_and_errors(multi):
    moments, names, qs = cq.hea_circuit(10, 2)
    prog = cq.serialize(moments)
    obs = cq.hea_observables(qs)
    B = 7
    vals = np.random.default_rng(3).uniform(0, 2, (B, len(names))).astype(np.float32)
    ref = ops.tfq_simulate_expectation([prog] * B, names, vals, [obs] * B, device=0)
    job = ops.DeviceJob("expectation", [prog] * B, names, vals, [obs] * B, device=multi)
    job.run()
    np.testing.assert_array_equal(job.fetch(), ref)
    job.close()
    down = np.ones((B, 4), np.float32)
    gref = ops.tfq_adj_grad([prog] * B, names, vals, [obs] * B, down, device=0)
    job = ops.DeviceJob("adjoint", [prog] * B, names, vals, [obs] * B, down, device=multi)
    job.run()
    np.testing.assert_allclose(job.fetch(), gref, atol=1e-6, rtol=1e-6)
    job.close()
    # fewer rows than devices, and an empty batch
    e = ops.tfq_simulate_expectation([prog], names, vals[:1], [obs], device=multi)
    np.testing.assert_array_equal(e, ref[:1])
    e = ops.tfq_simulate_expectation([], names, vals[:0], np.empty((0, 4), dtype=object),
                                     device=multi)
    assert e.shape[0] == 0
    # errors keep the reference's text, whichever block raises them
    with pytest.raises(ops.InvalidArgumentError, match="do not match"):
        ops.tfq_simulate_expectation([prog] * B, names, vals[:3], [obs] * B, device=multi)
    bad = [prog] * (B - 1) + [b"\xff\xff not a proto"]
    with pytest.raises(ops.InvalidArgumentError, match="Unparseable proto"):
        ops.tfq_simulate_expectation(bad, names, vals, [obs] * B, device=multi)
    with pytest.raises(ops.InvalidArgumentError, match="Could not find symbol"):
        ops.tfq_adj_grad([prog] * B, names[:-1], vals[:, :-1], [obs] * B, down, device=multi)


def test_row_offset_reproduces_unsplit_sampling():
    """a batch split by the caller (one rank per GPU) must draw the
    uniforms of the unsplit batch for the same seed."""
    n_list = [6, 6, 6, 6]
    progs, _ = _ragged(5, n_list)
    vals = np.zeros((4, 2), np.float32)
    ctx = ops.get_context(0)
    full = ops.tfq_simulate_samples(progs, ["a", "b"], vals, [50], seed=42, device=0)
    try:
        ctx.set_row_offset(2)
        tail = ops.tfq_simulate_samples(progs[2:], ["a", "b"], vals[2:], [50], seed=42,
                                        device=0)
    finally:
        ctx.set_row_offset(0)
    np.testing.assert_array_equal(tail, full[2:])


def test_multi_matches_single_noisy_ops(multi):
    """The noisy trajectory ops (N2) through a multi-device context: same
    Philox streams (keyed by global row), same values."""
    q = [cq.grid(0, i) for i in range(4)]
    progs = []
    for k in range(5):
        m = cq.random_circuit(q, 5, 40 + k, symbols=("a", "b"))
        m.insert(2, [cq.depolarize(q[k % 4], 0.2), cq.amplitude_damp(q[(k + 1) % 4], 0.3)])
        progs.append(cq.serialize(m))
    vals = np.random.default_rng(4).uniform(0, 2, (5, 2)).astype(np.float32)
    sums = [[cq.pauli_sum([(1.0, [(x, "Z")]) for x in q])]] * 5
    ns = np.full((5, 1), 23, np.int32)
    a = ops.tfq_noisy_expectation(progs, ["a", "b"], vals, sums, ns, seed=8, device=multi)
    b = ops.tfq_noisy_expectation(progs, ["a", "b"], vals, sums, ns, seed=8, device=0)
    np.testing.assert_array_equal(a, b)
    sa = ops.tfq_noisy_samples(progs, ["a", "b"], vals, [12], seed=8, device=multi)
    sb = ops.tfq_noisy_samples(progs, ["a", "b"], vals, [12], seed=8, device=0)
    np.testing.assert_array_equal(sa, sb)
