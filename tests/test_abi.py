"""CPU-only checks of the drop-in boundary: the C-ABI library loads and
exports every symbol include/tfqb.h declares, the host half (wire decoder,
qubit resolution, gate builders, PauliSum lowering, pass planner) agrees with
the oracle, and the product path fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

from oracle import tfq_oracle as orc
from quantum_b200 import circuits as cq
from quantum_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(ops.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return ops.load_library()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "tfqb.h")).read()
    declared = set(re.findall(r"\b(tfqb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(ops.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tfqb_abi_version() == 4


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    q = cq.grid(0, 0)
    with pytest.raises(ops.BackendUnavailableError):
        ops.tfq_simulate_expectation(
            [cq.serialize([[cq.X(q)]])], [], np.zeros((1, 0), np.float32),
            [[cq.pauli_sum([(1.0, [(q, "Z")])])]], device=0)


# gate kinds of csrc/program.h -> oracle closed forms
_KINDS_1Q = {2: orc.mat_xpow, 3: orc.mat_ypow, 4: orc.mat_zpow, 5: orc.mat_hpow}
_KINDS_2Q = {6: orc.mat_xxpow, 7: orc.mat_yypow, 8: orc.mat_zzpow,
             9: orc.mat_czpow, 10: orc.mat_cxpow, 11: orc.mat_swappow,
             12: orc.mat_iswappow}


def test_gate_matrices_bit_identical_to_oracle(lib):
    rng = np.random.default_rng(0)
    F = np.float32
    for _ in range(25):
        e, s, gs = F(rng.uniform(-2, 2)), F(rng.uniform(0, 1)), F(rng.uniform(-1, 1))
        t = F(e * s)
        for kind, fn in {**_KINDS_1Q, **_KINDS_2Q}.items():
            a = ops.host_gate_matrix(kind, [e, s, gs])
            b = np.asarray(fn(t, gs), dtype=np.complex64)
            np.testing.assert_array_equal(a, b.reshape(a.shape), err_msg=str(kind))
        p, ps = F(rng.uniform(-1, 1)), F(rng.uniform(0, 1))
        a = ops.host_gate_matrix(13, [p, ps, e, s, gs])
        np.testing.assert_array_equal(
            a, np.asarray(orc.mat_phasedxpow(F(p * ps), t, gs), np.complex64))
        a = ops.host_gate_matrix(14, [e, s, p, ps])
        np.testing.assert_array_equal(
            a, np.asarray(orc.mat_fsim(t, F(p * ps)), np.complex64))
        a = ops.host_gate_matrix(15, [p, ps, e, s])
        np.testing.assert_array_equal(
            a, np.asarray(orc.mat_phasediswappow(F(p * ps), t), np.complex64))


def test_gradient_gate_goldens_through_abi(lib):
    """adj_util_test.cc:355-398: Y^0.125 and XX^0.001 gradient gates."""
    g = ops.host_gate_matrix(3, [0.125, 1.0, 0.0], grad_param=0)
    ref = (np.asarray(orc.mat_ypow(np.float32(np.float32(0.125) + orc.GRAD_EPS)), np.complex64) -
           np.asarray(orc.mat_ypow(np.float32(np.float32(0.125) - orc.GRAD_EPS)), np.complex64))
    ref = (ref.view(np.float32) * np.float32(0.5 / 5e-3)).view(np.complex64)
    np.testing.assert_allclose(g, ref.reshape(2, 2), atol=2e-5)
    # reference golden values (adj_util_test.cc:362-374, tol 1e-4)
    gold = np.array([[-0.60111 + 1.45122j, -1.45122 - 0.60111j],
                     [1.45122 + 0.60111j, -0.60111 + 1.45122j]])
    np.testing.assert_allclose(g, gold, atol=1e-4)
    # XX^0.001 (adj_util_test.cc:377-404, tol 1e-4)
    g = ops.host_gate_matrix(6, [0.001, 1.0, 0.0], grad_param=0)
    d, o = -0.004934 + 1.57078j, 0.004934 - 1.57078j
    gold = np.array([[d, 0, 0, o], [0, d, o, 0], [0, o, d, 0], [o, 0, 0, d]])
    np.testing.assert_allclose(g, gold, atol=1e-4)


def test_plan_describes_qubit_mapping_and_controls(lib):
    """circuit_parser_qsim_test.cc:110-266: proto qubit i of n -> bit n-1-i,
    controls reversed the same way; program_resolution.cc:121-123 ordering
    (grid qubits by (row, col), line qubits last)."""
    qs = [cq.grid(0, 1), cq.grid(0, 0), cq.line(3), cq.grid(1, 0)]
    op = cq.X(qs[0], 0.5).controlled_by([qs[2], qs[3]], [1, 0])
    d = ops.host_describe_plan(cq.serialize([[op], [cq.CNOT(qs[1], qs[3])]]))
    # sorted order: 0_0, 0_1, 1_0, line 3  -> indices 0,1,2,3 ; bit = 3 - idx
    assert d["n"] == 4
    g0, g1 = d["gates"]
    assert g0["bits"] == [2]
    assert g0["cmask"] == (1 << 0) | (1 << 1)     # line3 -> bit0, 1_0 -> bit1
    assert g0["cbits"] == (1 << 0)                # line3 must be 1, 1_0 must be 0
    assert g1["bits"] == [3, 1]
    assert d["n_alloc"] == 5 and len(d["passes"]) >= 1


def test_pauli_sum_lowering(lib):
    qs = [cq.grid(0, i) for i in range(3)]
    prog = cq.serialize([[cq.H(q) for q in qs]])
    ps = cq.pauli_sum([(0.5, [(qs[0], "X"), (qs[2], "Y")]), (2.0, []),
                       (-1.0, [(qs[1], "Z")])])
    d = ops.host_describe_pauli_sum(prog, ps)
    t0, t1, t2 = d["terms"]
    assert (t0["x"], t0["z"], t0["phase"]) == (0b101, 0b001, 1)
    assert t0["parity_mask"] == 0b101 and abs(t0["coeff"] - 0.5) < 1e-7
    assert t1["identity"] == 1
    assert (t2["x"], t2["z"], t2["phase"]) == (0, 0b010, 0)
    with pytest.raises(ops.InvalidArgumentError, match="qubits not found in circuit"):
        ops.host_describe_pauli_sum(prog, cq.pauli_sum([(1.0, [(cq.grid(9, 9), "Z")])]))


def test_text_format_programs_parse(lib):
    """parse_context.cc:41-56: binary first, then text format (the stock
    random-circuit benchmark sends text, benchmark_random_circuit.py:103)."""
    qs = [cq.grid(0, i) for i in range(3)]
    m = cq.random_circuit(qs, 5, 3, controls=True, symbols=("a",))
    a = ops.host_describe_plan(cq.serialize(m), ["a"])
    b = ops.host_describe_plan(cq.serialize_text(m), ["a"])
    assert a == b
    with pytest.raises(ops.InvalidArgumentError, match="Unparseable proto"):
        ops.host_describe_plan(b"\xff\xfe not a proto")
    with pytest.raises(ops.InvalidArgumentError, match="Could not find symbol"):
        ops.host_describe_plan(cq.serialize(m), ["zzz"])


def test_planner_invariants(lib):
    """Every non-identity gate is scheduled exactly once (forward) and the
    adjoint plan carries one gradient slot per (gate, symbol)."""
    for n, seed in ((5, 1), (13, 2), (17, 3), (22, 4)):
        qs = [cq.grid(0, i) for i in range(n)]
        m = cq.random_circuit(qs, 12, seed, controls=True, symbols=("a", "b"))
        d = ops.host_describe_plan(cq.serialize(m), ["a", "b"])
        n_gates = sum(1 for g in d["gates"] if g["kind"] > 1)
        # every gate exactly once (+ identity place-holders of the product-
        # state init for qubits whose first gate is not a 1-qubit gate)
        assert d["n_factors"] - d["init_identity_bits"] == n_gates
        assert d["n_ops"] <= n_gates              # fusion only merges
        assert sum(p["ops"] for p in d["passes"]) == d["n_ops"]
        for p in d["passes"]:
            assert p["tile"][:4] == [0, 1, 2, 3] or n < 4
            assert len(set(p["tile"])) == len(p["tile"]) == min(12, max(n, 5))
        da = ops.host_describe_plan(cq.serialize(m), ["a", "b"], adjoint=True)
        n_sym = sum(len(g["syms"]) for g in d["gates"] if g["kind"] > 1)
        assert len(da["grad_slots"]) == n_sym
        # un-controlled gates with one symbol are one fused adjoint op
        # (psi <- G'psi, gradient, lam <- G'lam); otherwise dag, grads, dag
        expect = 0
        for g in d["gates"]:
            if g["kind"] <= 1:
                continue
            k = len(g["syms"])
            expect += 1 if (k == 0 or (k == 1 and g["cmask"] == 0)) else 2 + k
        # macro-ops merge commuting ops of a round into one dispatch
        assert da["n_ops"] + da["macro_merged"] == expect


def test_workload_plans_are_few_passes(lib):
    moments, names, _ = cq.hea_circuit(20, 4)
    d = ops.host_describe_plan(cq.serialize(moments), names)
    assert len(d["passes"]) <= 4, "C2 forward should stay within 4 HBM passes"


def test_sharded_plan_invariants(lib):
    """Stage list of the sharded-state planner: starts with a gate segment,
    exchanges separate segments, the final qubit layout is a permutation, and
    every gate is applied exactly once (SWAPs that move qubits to the
    exchange positions are extra)."""
    for n, world, seed in ((14, 2, 1), (16, 4, 2), (18, 8, 3)):
        qs = [cq.grid(0, i) for i in range(n)]
        m = cq.random_circuit(qs, 10, seed, controls=True)
        sums = [cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs]),
                cq.pauli_sum([(1.0, [(q, "X")]) for q in qs])]
        d = ops.host_describe_sharded(cq.serialize(m), [], sums, world)
        flat = ops.host_describe_plan(cq.serialize(m))
        assert d["n_local"] == n - int(np.log2(world))
        assert d["stages"][0]["kind"] == 0
        assert sorted(d["final_phys"]) == list(range(n))
        kinds = [s["kind"] for s in d["stages"]]
        assert kinds.count(1) == d["n_exchanges"]
        assert 2 in kinds and kinds[-1] == 2
        n_factors = sum(s["factors"] for s in d["stages"] if s["kind"] == 0)
        assert n_factors >= flat["n_factors"] - flat["init_identity_bits"]
        # all X terms end up evaluated: the last expectation stage defers none
        last = [s for s in d["stages"] if s["kind"] == 2][-1]
        assert last["deferred"] == 0
    # tiles a specialised CTA walks: the first pass of a segment that follows
    # a qubit swap gathers from the peers and keeps 2, local passes 8 (32 from 2^28 amplitudes on)
    qs = [cq.grid(0, i) for i in range(30)]
    m = cq.random_circuit(qs, 8, 5)
    d = ops.host_describe_sharded(cq.serialize(m), [],
                                  [cq.pauli_sum([(1.0, [(q, "Z")]) for q in qs])], 4)
    segs = [s for s in d["stages"] if s["kind"] == 0]
    assert len(segs) >= 2
    local = 32 if d["n_local"] >= 28 else 8      # large shards: 32 (shared 2 MB pages)
    assert d["n_local"] == 28
    for k, sg in enumerate(segs):
        want0 = 2 if sg["after_exchange"] else local
        assert sg["tiles_per_cta"][0] == want0, (k, sg)
        assert all(t == local for t in sg["tiles_per_cta"][1:])
    assert segs[0]["after_exchange"] == 0 and segs[1]["after_exchange"] == 1


def _nvrtc_compile(src):
    """Compile CUDA C++ text for sm_100a with the toolkit's NVRTC (no GPU
    needed); returns (rc, log)."""
    import ctypes
    rtc = None
    for name in ("libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12",
                 "/usr/local/cuda/lib64/libnvrtc.so"):
        try:
            rtc = ctypes.CDLL(name)
            break
        except OSError:
            continue
    if rtc is None:
        pytest.skip("libnvrtc not present")
    prog = ctypes.c_void_p()
    assert rtc.nvrtcCreateProgram(ctypes.byref(prog), src.encode(), b"k.cu", 0,
                                  None, None) == 0
    opts = (ctypes.c_char_p * 2)(b"--gpu-architecture=sm_100a", b"-std=c++17")
    rc = rtc.nvrtcCompileProgram(prog, 2, opts)
    n = ctypes.c_size_t()
    rtc.nvrtcGetProgramLogSize(prog, ctypes.byref(n))
    log = ctypes.create_string_buffer(n.value)
    rtc.nvrtcGetProgramLog(prog, log)
    rtc.nvrtcDestroyProgram(ctypes.byref(prog))
    return rc, log.value.decode()


def test_specialised_pass_kernels_compile_for_sm100a(lib):
    """csrc/jit.cc: every pass of the HEA and TFI workloads (forward and
    adjoint plans) has a specialised kernel whose generated source NVRTC
    accepts for sm_100a; passes with controlled gates are left to the
    interpreted kernel."""
    moments, names, _ = cq.hea_circuit(13, 2)
    prog = cq.serialize(moments)
    for adjoint in (False, True):
        d = ops.host_describe_plan(prog, names, adjoint=adjoint)
        assert d["passes"]
        for p in range(len(d["passes"])):
            src = ops.host_jit_source(prog, names, adjoint=adjoint, pass_index=p)
            assert "tfqb_jit_pass" in src and "g1_packed" in src
            rc, log = _nvrtc_compile(src)
            assert rc == 0, log[:2000]
    # a random circuit: the kernels of its un-controlled passes compile too
    qs = [cq.grid(0, i) for i in range(14)]
    m = cq.random_circuit(qs, 8, 5, controls=False, symbols=("a", "b"))
    prog = cq.serialize(m)
    n_src = 0
    for adjoint in (False, True):
        d = ops.host_describe_plan(prog, ["a", "b"], adjoint=adjoint)
        for p in range(len(d["passes"])):
            src = ops.host_jit_source(prog, ["a", "b"], adjoint=adjoint, pass_index=p)
            if src:
                n_src += 1
                rc, log = _nvrtc_compile(src)
                assert rc == 0, log[:2000]
    assert n_src > 0
    # jobs that cannot see a global phase get "phased real" gates: Y^a then
    # Z^b is diag(1, q) R, Z^b then Y^a is R diag(1, q); X^a is general
    qs = [cq.grid(0, i) for i in range(13)]
    lead = [[cq.H(q) for q in qs], [cq.CZ(qs[i], qs[i + 1]) for i in range(0, 12, 2)]]
    for first, second, want in ((cq.Y, cq.Z, "g1_rowreal_lift<"), (cq.Z, cq.Y, "g1_colreal_lift<"),
                                (cq.X, cq.Z, "g1_packed<")):
        m = lead + [[first(q, "a") for q in qs], [second(q, "b") for q in qs],
                    [cq.CZ(qs[i], qs[i + 1]) for i in range(1, 12, 2)]]
        prog = cq.serialize(m)
        src = ops.host_jit_source(prog, ["a", "b"], pass_index=0, phase_free=True)
        assert want in src, want
        assert "phased_real_setup(s_mat" in src or want == "g1_packed<"
        exact = ops.host_jit_source(prog, ["a", "b"], pass_index=0)
        assert "g1_rowreal<" not in exact and "g1_colreal<" not in exact
        rc, log = _nvrtc_compile(src)
        assert rc == 0, log[:2000]
    # a literal H is a REFLECTION times a phase: real matrix forms, never shears
    m = lead + [[cq.H(q) for q in qs], [cq.Z(q, "b") for q in qs],
                [cq.CZ(qs[i], qs[i + 1]) for i in range(1, 12, 2)],
                [cq.Y(q, "a") for q in qs], [cq.H(q) for q in qs]]
    src = ops.host_jit_source(cq.serialize(m), ["a", "b"], pass_index=0, phase_free=True)
    assert "g1_rowreal<" in src and "g1_real<" in src
    assert "g1_rowreal_lift<" not in src and "g1_real_lift<" not in src
    rc, log = _nvrtc_compile(src)
    assert rc == 0, log[:2000]
    # X^a alone is a phase times [[c, -i s], [-i s, c]]: 2 packed FMAs per amplitude,
    # forward and (with the dropped phase moved into the gradient gate) adjoint
    m = lead + [[cq.X(q, "a") for q in qs], [cq.CZ(qs[i], qs[i + 1]) for i in range(1, 12, 2)],
                [cq.Y(q, "b") for q in qs]]
    prog = cq.serialize(m)
    src = ops.host_jit_source(prog, ["a", "b"], pass_index=0, phase_free=True)
    assert "g1_ximag_lift<" in src and "g1_real_lift<" in src and "phased_ximag_setup(" in src
    rc, log = _nvrtc_compile(src)
    assert rc == 0, log[:2000]
    src = ops.host_jit_source(prog, ["a", "b"], adjoint=True, pass_index=0)
    assert "adj1_ximag_lift<" in src and "adj1_real_lift<" in src
    rc, log = _nvrtc_compile(src)
    assert rc == 0, log[:2000]
    # fewer than 12 qubits: no full tile, nothing to specialise
    moments, names, _ = cq.hea_circuit(8, 2)
    assert ops.host_jit_source(cq.serialize(moments), names) == ""
    # PauliSum expectation passes of the C2 observables
    moments, names, qs = cq.hea_circuit(15, 2)
    prog = cq.serialize(moments)
    n_src = 0
    for p in range(4):
        src = ops.host_jit_expect_source(prog, cq.hea_observables(qs), p)
        if src:
            n_src += 1
            assert "tfqb_jit_expect" in src
            rc, log = _nvrtc_compile(src)
            assert rc == 0, log[:2000]
    assert n_src >= 1
    src = ops.host_jit_expect_source(prog, cq.hea_observables(qs), 1000)
    assert "tfqb_jit_accum" in src
    rc, log = _nvrtc_compile(src)
    assert rc == 0, log[:2000]


def test_malformed_qubit_ids_are_errors_not_crashes(lib):
    """an empty qubit id, or one holding ',', used to throw
    std::out_of_range across the C ABI and abort the process."""
    for bad in ("", "0_0,0_1", "x_y"):
        p = cq.to_program([[cq.X(cq.grid(0, 0), 0.5), cq.X(cq.grid(0, 5), 0.5)]])
        p.circuit.moments[0].operations[0].qubits[0].id = bad
        with pytest.raises(ops.InvalidArgumentError, match="Unable to parse qubit"):
            ops.host_describe_plan(p.SerializeToString())


def test_jit_pending_counter(lib):
    """tfqb_jit_pending / ops.wait_for_jit: no compilation runs in a process
    that has not asked for one (bench.py waits on this before host-timed legs)."""
    assert ops.jit_pending() == 0
    assert ops.wait_for_jit(limit_s=1.0) < 1.0
