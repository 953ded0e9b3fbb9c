"""The TensorFlow shim (quantum_b200/csrc/tf_ops/tfq_b200_ops.cc: the
DEVICE_GPU registrations of the five ops + the two inner-product ops) compiled
against a minimal stand-in of TensorFlow's op-kernel API (tests/tf_stub/;
TensorFlow itself is not in this image) and linked with libtfqb.so.

CPU: it compiles and links, and registers every op for DEVICE_GPU with all
arguments in host memory.  GPU: its Compute() methods run on host tensors and
return what the C ABI returns through ctypes.
"""
import os
import subprocess

import numpy as np
import pytest

from quantum_b200 import circuits as cq
from quantum_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "tf_stub")
SHIM = os.path.join(ROOT, "quantum_b200", "csrc", "tf_ops", "tfq_b200_ops.cc")
OPS = ["TfqSimulateExpectation", "TfqSimulateSampledExpectation", "TfqSimulateSamples",
       "TfqSimulateState", "TfqAdjointGradient", "TfqInnerProduct", "TfqInnerProductGrad",
       "TfqNoisyExpectation", "TfqNoisySampledExpectation", "TfqNoisySamples",
       "TfqCalculateUnitary"]


def _build(tmp):
    exe = os.path.join(tmp, "tf_shim_harness")
    lib_dir = os.path.join(ROOT, "quantum_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I", STUB, "-I", os.path.join(ROOT, "include"),
           SHIM, os.path.join(STUB, "harness.cc"), "-o", exe,
           "-L", lib_dir, "-l:libtfqb.so", "-Wl,-rpath," + lib_dir,
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart", "-ldl",
           "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return exe


def _write_inputs(path, progs, names, vals, sums, down):
    def s(b):
        return str(len(b)).encode() + b"\n" + b + b"\n"
    B, P, M = len(progs), len(names), len(sums[0])
    with open(path, "wb") as f:
        f.write(("%d %d %d\n" % (B, P, M)).encode())
        for p in progs:
            f.write(s(p))
        for n in names:
            f.write(s(n.encode()))
        f.write((" ".join("%.9g" % v for v in vals.reshape(-1)) + "\n").encode())
        for row in sums:
            for ps in row:
                f.write(s(ps))
        f.write((" ".join("%.9g" % v for v in down.reshape(-1)) + "\n").encode())


def _case():
    moments, names, qs = cq.hea_circuit(6, 2)
    prog = cq.serialize(moments)
    obs = cq.hea_observables(qs)
    B = 3
    vals = np.random.default_rng(5).uniform(0, 2, (B, len(names))).astype(np.float32)
    down = np.random.default_rng(6).normal(size=(B, len(obs))).astype(np.float32)
    return [prog] * B, names, vals, [obs] * B, down


def test_shim_compiles_links_and_registers_gpu_kernels(tmp_path):
    ops.load_library()            # libtfqb.so must exist (no CPU fallback)
    exe = _build(str(tmp_path))
    assert os.path.exists(exe)
    src = open(SHIM).read()
    for op in OPS:
        assert 'TFQB_GPU_KERNEL("%s"' % op in src


@pytest.mark.gpu
def test_shim_compute_matches_c_abi(tmp_path):
    exe = _build(str(tmp_path))
    progs, names, vals, sums, down = _case()
    inp = os.path.join(str(tmp_path), "inputs.bin")
    _write_inputs(inp, progs, names, vals, sums, down)
    res = subprocess.run([exe, inp], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = {l.split(" ", 1)[0]: l.split(" ", 1)[1] for l in res.stdout.strip().splitlines()}
    for op in OPS:
        assert op + ":GPU:" in lines["registered"]
    e = np.array(lines["TfqSimulateExpectation"].split(), np.float32).reshape(len(progs), -1)
    g = np.array(lines["TfqAdjointGradient"].split(), np.float32).reshape(len(progs), -1)
    np.testing.assert_array_equal(e, ops.tfq_simulate_expectation(progs, names, vals, sums))
    np.testing.assert_allclose(g, ops.tfq_adj_grad(progs, names, vals, sums, down),
                               atol=1e-6, rtol=1e-6)
    assert lines["rank_error"].startswith("3 symbol_values must be rank 2. Got rank 1.")
