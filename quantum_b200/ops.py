"""Host-side mirror of TFQ's circuit-execution op wrappers, bound to the
B200 backend through its C ABI (include/tfqb.h) with ctypes.

Same names, argument meaning and error behaviour as the reference wrappers:
  tfq_simulate_expectation / tfq_simulate_state / tfq_simulate_samples /
  tfq_simulate_sampled_expectation
      tensorflow_quantum/core/ops/tfq_simulate_ops.py:23-135
  tfq_adj_grad
      tensorflow_quantum/core/ops/tfq_adj_grad_op.py:22-48
Inputs are what the TF ops receive: `programs` / `pauli_sums` are arrays of
serialized `tfq.proto.Program` / `tfq.proto.PauliSum` strings, `symbol_names`
an array of strings, `symbol_values` a [batch, n_symbols] float array.  Rank
errors that TF's shape functions / GetProgramsAndNumQubits raise are raised
here with the same text (parse_context.cc:70-73,263-266,301-303,313-316).

There is no CPU path in this module: if libtfqb.so is not built, or no CUDA
device is visible, every op raises.
"""
from __future__ import annotations

import ctypes
import json
import os
import threading
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TFQB_LIB") or os.path.join(_HERE, "libtfqb.so")

TFQB_OK = 0
TFQB_INVALID_ARGUMENT = 3
TFQB_RESOURCE_EXHAUSTED = 8
TFQB_INTERNAL = 13
TFQB_UNAVAILABLE = 14


class InvalidArgumentError(ValueError):
    """Stands in for tf.errors.InvalidArgumentError."""


class BackendUnavailableError(RuntimeError):
    """libtfqb.so missing or no CUDA device (there is no CPU fallback)."""


class _Strings(ctypes.Structure):
    _fields_ = [("data", ctypes.POINTER(ctypes.c_char_p)),
                ("size", ctypes.POINTER(ctypes.c_size_t))]


class _StringList(ctypes.Structure):
    _fields_ = [("data", ctypes.POINTER(ctypes.c_void_p)),
                ("size", ctypes.POINTER(ctypes.c_size_t)),
                ("count", ctypes.c_size_t)]

    def take(self):
        """Copy the strings out and hand the list back to the library."""
        out = [ctypes.string_at(self.data[i], self.size[i]) for i in range(self.count)]
        load_library().tfqb_free_string_list(ctypes.byref(self))
        return out


class _CircuitInputs(ctypes.Structure):
    _fields_ = [("programs", _Strings), ("batch", ctypes.c_int),
                ("symbol_names", _Strings), ("n_symbols", ctypes.c_int),
                ("symbol_values", ctypes.POINTER(ctypes.c_float)),
                ("symbol_rows", ctypes.c_int)]


class Profile(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_int64),
                ("gate_pass_launches", ctypes.c_int64),
                ("adjoint_pass_launches", ctypes.c_int64),
                ("gate_pass_ms", ctypes.c_double),
                ("adjoint_pass_ms", ctypes.c_double),
                ("gate_pass_bytes", ctypes.c_double),
                ("adjoint_pass_bytes", ctypes.c_double),
                ("h2d_bytes", ctypes.c_int64),
                ("d2h_bytes", ctypes.c_int64),
                ("expectation_launches", ctypes.c_int64),
                ("expectation_ms", ctypes.c_double),
                ("expectation_bytes", ctypes.c_double),
                ("jit_kernels", ctypes.c_int64),
                ("jit_pass_launches", ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


PEER_HANDLE_BYTES = 256


class ExchangeStats(ctypes.Structure):
    _fields_ = [("n_qubits", ctypes.c_int), ("n_local", ctypes.c_int),
                ("exchanges", ctypes.c_int), ("fused_exchanges", ctypes.c_int),
                ("gate_passes", ctypes.c_int),
                ("expectation_passes", ctypes.c_int),
                ("shard_bytes", ctypes.c_double),
                ("bytes_received_per_exchange", ctypes.c_double),
                ("wait_ms", ctypes.c_double), ("pull_ms", ctypes.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/tfqb.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "tfqb_abi_version", "tfqb_create", "tfqb_create_multi", "tfqb_device_count",
    "tfqb_set_row_offset", "tfqb_destroy", "tfqb_last_error",
    "tfqb_set_memory_budget", "tfqb_trim", "tfqb_jit_compile_seconds",
    "tfqb_simulate_expectation",
    "tfqb_simulate_sampled_expectation", "tfqb_simulate_samples_prepare",
    "tfqb_simulate_samples_run", "tfqb_simulate_state_prepare",
    "tfqb_simulate_state_run", "tfqb_adjoint_gradient",
    "tfqb_expectation_prepare", "tfqb_adjoint_prepare", "tfqb_job_run_device",
    "tfqb_job_fetch", "tfqb_job_free", "tfqb_sync", "tfqb_stream",
    "tfqb_profile_enable", "tfqb_profile_reset", "tfqb_profile_read",
    "tfqb_calculate_unitary_prepare", "tfqb_calculate_unitary_run",
    "tfqb_noisy_expectation", "tfqb_noisy_sampled_expectation",
    "tfqb_noisy_samples_prepare", "tfqb_noisy_samples_run",
    "tfqb_inner_product", "tfqb_inner_product_grad", "tfqb_sharded_prepare", "tfqb_sharded_stage_kind", "tfqb_sharded_run_stage",
    "tfqb_sharded_buffers", "tfqb_sharded_partials", "tfqb_sharded_finish",
    "tfqb_sharded_export", "tfqb_sharded_connect", "tfqb_sharded_enqueue",
    "tfqb_sharded_result", "tfqb_sharded_stats", "tfqb_sharded_sample",
    "tfqb_host_gate_matrix", "tfqb_host_describe_plan",
    "tfqb_host_describe_pauli_sum", "tfqb_host_describe_sharded",
    "tfqb_host_jit_source", "tfqb_host_jit_expect_source", "tfqb_free_string",
    "tfqb_ps_decompose", "tfqb_ps_symbol_replace", "tfqb_ps_weights_from_symbols",
    "tfqb_free_string_list", "tfqb_free_floats", "tfqb_jit_pending",
]

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen quantum_b200/libtfqb.so (built in-tree by build.sh)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise BackendUnavailableError(
                "%s not found: run ./build.sh (or __graft_entry__.build()). "
                "The B200 backend has no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        lib.tfqb_last_error.restype = ctypes.c_char_p
        lib.tfqb_create.argtypes = [ci, ctypes.POINTER(vp)]
        lib.tfqb_create_multi.argtypes = [ctypes.POINTER(ci), ci, ctypes.POINTER(vp)]
        lib.tfqb_device_count.argtypes = [vp]
        lib.tfqb_set_row_offset.argtypes = [vp, ctypes.c_int64]
        lib.tfqb_destroy.argtypes = [vp]
        lib.tfqb_destroy.restype = None
        lib.tfqb_set_memory_budget.argtypes = [vp, ctypes.c_size_t]
        lib.tfqb_trim.argtypes = [vp]
        lib.tfqb_jit_compile_seconds.restype = ctypes.c_double
        pin = ctypes.POINTER(_CircuitInputs)
        fp = ctypes.POINTER(ctypes.c_float)
        lib.tfqb_simulate_expectation.argtypes = [vp, pin, _Strings, ci, ci, fp]
        lib.tfqb_simulate_sampled_expectation.argtypes = [
            vp, pin, _Strings, ci, ci, ctypes.POINTER(ctypes.c_int32), ci, ci,
            ctypes.c_uint64, ctypes.POINTER(ctypes.c_double), ci, ci, fp]
        lib.tfqb_simulate_samples_prepare.argtypes = [
            vp, pin, ci, ctypes.POINTER(vp), ctypes.POINTER(ci)]
        lib.tfqb_simulate_samples_run.argtypes = [
            vp, ctypes.c_uint64, ctypes.POINTER(ctypes.c_double),
            ctypes.POINTER(ctypes.c_int8)]
        lib.tfqb_simulate_state_prepare.argtypes = [
            vp, pin, ctypes.POINTER(vp), ctypes.POINTER(ci)]
        lib.tfqb_simulate_state_run.argtypes = [vp, fp]
        lib.tfqb_adjoint_gradient.argtypes = [vp, pin, _Strings, ci, ci, fp,
                                              ci, ci, fp]
        lib.tfqb_expectation_prepare.argtypes = [vp, pin, _Strings, ci, ci,
                                                 ctypes.POINTER(vp)]
        lib.tfqb_adjoint_prepare.argtypes = [vp, pin, _Strings, ci, ci, fp, ci,
                                             ci, ctypes.POINTER(vp)]
        lib.tfqb_calculate_unitary_prepare.argtypes = [
            vp, pin, ctypes.POINTER(vp), ctypes.POINTER(ci)]
        lib.tfqb_calculate_unitary_run.argtypes = [vp, fp]
        noisy_args = [vp, pin, _Strings, ci, ci, ctypes.POINTER(ctypes.c_int32), ci, ci,
                      ctypes.c_uint64, fp, ci, ci, fp]
        lib.tfqb_noisy_expectation.argtypes = noisy_args
        lib.tfqb_noisy_sampled_expectation.argtypes = noisy_args
        lib.tfqb_noisy_samples_prepare.argtypes = [
            vp, pin, ci, ctypes.POINTER(vp), ctypes.POINTER(ci)]
        lib.tfqb_noisy_samples_run.argtypes = [
            vp, ctypes.c_uint64, fp, ci, ctypes.POINTER(ctypes.c_double),
            ctypes.POINTER(ctypes.c_int8)]
        lib.tfqb_inner_product.argtypes = [vp, pin, _Strings, ci, ci, fp]
        lib.tfqb_inner_product_grad.argtypes = [vp, pin, _Strings, ci, ci, fp, ci, ci, fp]
        lib.tfqb_sharded_prepare.argtypes = [
            vp, pin, _Strings, ci, ci, ci, ctypes.POINTER(vp),
            ctypes.POINTER(ci), ctypes.POINTER(ci)]
        lib.tfqb_sharded_stage_kind.argtypes = [vp, ci]
        lib.tfqb_sharded_run_stage.argtypes = [vp, ci]
        lib.tfqb_sharded_buffers.argtypes = [
            vp, ctypes.POINTER(vp), ctypes.POINTER(vp),
            ctypes.POINTER(ctypes.c_size_t)]
        lib.tfqb_sharded_partials.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
        lib.tfqb_sharded_finish.argtypes = [vp, ctypes.POINTER(ctypes.c_double), fp]
        lib.tfqb_sharded_export.argtypes = [vp, ctypes.c_char_p]
        lib.tfqb_sharded_connect.argtypes = [vp, ctypes.c_char_p, ci]
        lib.tfqb_sharded_enqueue.argtypes = [vp]
        lib.tfqb_sharded_result.argtypes = [vp, fp]
        lib.tfqb_sharded_stats.argtypes = [vp, ctypes.POINTER(ExchangeStats)]
        lib.tfqb_sharded_sample.argtypes = [
            vp, ci, ctypes.c_uint64, ctypes.POINTER(ctypes.c_double),
            ctypes.POINTER(ctypes.c_int8), ctypes.POINTER(ctypes.c_int32)]
        lib.tfqb_job_run_device.argtypes = [vp]
        lib.tfqb_job_fetch.argtypes = [vp, fp]
        lib.tfqb_job_free.argtypes = [vp]
        lib.tfqb_job_free.restype = None
        lib.tfqb_sync.argtypes = [vp]
        lib.tfqb_stream.argtypes = [vp]
        lib.tfqb_stream.restype = vp
        lib.tfqb_profile_enable.argtypes = [vp, ci]
        lib.tfqb_profile_reset.argtypes = [vp]
        lib.tfqb_profile_read.argtypes = [vp, ctypes.POINTER(Profile)]
        lib.tfqb_host_gate_matrix.argtypes = [ci, fp, ci, ci, fp]
        lib.tfqb_host_describe_plan.argtypes = [
            ctypes.c_char_p, ctypes.c_size_t, _Strings, ci, ci,
            ctypes.POINTER(ctypes.c_char_p)]
        lib.tfqb_host_describe_pauli_sum.argtypes = [
            ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
            ctypes.POINTER(ctypes.c_char_p)]
        lib.tfqb_host_describe_sharded.argtypes = [
            ctypes.c_char_p, ctypes.c_size_t, _Strings, ci, _Strings, ci, ci,
            ctypes.POINTER(ctypes.c_char_p)]
        lib.tfqb_host_jit_source.argtypes = [
            ctypes.c_char_p, ctypes.c_size_t, _Strings, ci, ci, ci,
            ctypes.POINTER(ctypes.c_char_p)]
        lib.tfqb_host_jit_expect_source.argtypes = [
            ctypes.c_char_p, ctypes.c_size_t, _Strings, ci, ci,
            ctypes.POINTER(ctypes.c_char_p)]
        lib.tfqb_free_string.argtypes = [ctypes.c_void_p]
        lib.tfqb_free_string.restype = None
        sl = ctypes.POINTER(_StringList)
        lib.tfqb_ps_decompose.argtypes = [_Strings, ci, sl]
        lib.tfqb_ps_symbol_replace.argtypes = [_Strings, ci, _Strings, ci, _Strings, ci, sl,
                                               ctypes.POINTER(ci)]
        lib.tfqb_ps_weights_from_symbols.argtypes = [
            _Strings, ci, _Strings, ci, ctypes.POINTER(fp), ctypes.POINTER(ci)]
        lib.tfqb_free_string_list.argtypes = [sl]
        lib.tfqb_free_string_list.restype = None
        lib.tfqb_free_floats.argtypes = [ctypes.c_void_p]
        lib.tfqb_free_floats.restype = None
        _lib = lib
        return lib


def _check(rc):
    if rc == TFQB_OK:
        return
    msg = load_library().tfqb_last_error().decode("utf-8", "replace")
    if rc == TFQB_INVALID_ARGUMENT:
        raise InvalidArgumentError(msg)
    if rc == TFQB_UNAVAILABLE:
        raise BackendUnavailableError(msg)
    if rc == TFQB_RESOURCE_EXHAUSTED:
        raise MemoryError(msg)
    raise RuntimeError("tfqb error %d: %s" % (rc, msg))


# --------------------------------------------------------------------------
# contexts: one per (process, GPU)
# --------------------------------------------------------------------------
class Context:
    """`device`: one CUDA ordinal, or a sequence of ordinals for a context
    that spreads the rows of every op call over several GPUs
    (tfqb_create_multi)."""

    def __init__(self, device):
        lib = load_library()
        self.device = device
        self._h = ctypes.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (ctypes.c_int * len(device))(*[int(d) for d in device])
            _check(lib.tfqb_create_multi(ids, len(device), ctypes.byref(self._h)))
        else:
            _check(lib.tfqb_create(device, ctypes.byref(self._h)))

    def device_count(self) -> int:
        return int(load_library().tfqb_device_count(self._h))

    def set_row_offset(self, first_row: int):
        """Global index of the first row this context is given: keeps the
        sampling ops' Philox streams those of the unsplit batch."""
        _check(load_library().tfqb_set_row_offset(self._h, int(first_row)))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            load_library().tfqb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def sync(self):
        _check(load_library().tfqb_sync(self._h))

    def stream(self) -> int:
        return int(load_library().tfqb_stream(self._h) or 0)

    def set_memory_budget(self, nbytes: int):
        _check(load_library().tfqb_set_memory_budget(self._h, nbytes))

    def trim(self):
        """Return the cached device memory to CUDA."""
        _check(load_library().tfqb_trim(self._h))

    def profile_enable(self, on: bool):
        _check(load_library().tfqb_profile_enable(self._h, 1 if on else 0))

    def profile_reset(self):
        _check(load_library().tfqb_profile_reset(self._h))

    def profile_read(self) -> dict:
        p = Profile()
        _check(load_library().tfqb_profile_read(self._h, ctypes.byref(p)))
        return p.as_dict()


_contexts = {}
_ctx_lock = threading.Lock()


def jit_compile_seconds() -> float:
    return float(load_library().tfqb_jit_compile_seconds())


def jit_pending() -> int:
    """Kernel compilations still running on background host threads."""
    return int(load_library().tfqb_jit_pending())


def wait_for_jit(limit_s: float = 120.0) -> float:
    """Block until no kernel compilation runs in the background; returns the
    seconds waited (bench legs that time HOST work call this first)."""
    import time
    t0 = time.perf_counter()
    while jit_pending() > 0 and time.perf_counter() - t0 < limit_s:
        time.sleep(0.01)
    return time.perf_counter() - t0


def default_device() -> int:
    for k in ("TFQB_DEVICE", "LOCAL_RANK"):
        if os.environ.get(k, "") != "":
            return int(os.environ[k])
    return 0


def get_context(device=None) -> Context:
    """`device`: ordinal, sequence of ordinals (one context over several
    GPUs), or None for TFQB_DEVICE / LOCAL_RANK / 0."""
    if device is None:
        device = default_device()
    if isinstance(device, list):
        device = tuple(device)
    with _ctx_lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = Context(device)
            _contexts[device] = ctx
        return ctx


# --------------------------------------------------------------------------
# marshalling
# --------------------------------------------------------------------------
def _as_bytes(x) -> bytes:
    if isinstance(x, bytes):
        return x
    if isinstance(x, str):
        return x.encode()
    if isinstance(x, (np.bytes_, np.str_)):
        return bytes(x) if isinstance(x, np.bytes_) else str(x).encode()
    if hasattr(x, "SerializeToString"):
        return x.SerializeToString()
    return bytes(x)


def _rank(x) -> int:
    if isinstance(x, np.ndarray):
        return x.ndim
    if isinstance(x, (bytes, str)) or hasattr(x, "SerializeToString"):
        return 0
    if isinstance(x, (list, tuple)):
        return 1 + (_rank(x[0]) if len(x) else 0)
    return np.ndim(x)


class _StringPack:
    """Keeps the bytes objects and ctypes arrays of a string tensor alive."""

    def __init__(self, flat: Sequence[bytes], tile: int = 1):
        """`tile` > 1: the table is `flat` repeated `tile` times (a batch whose
        rows are one and the same list object)."""
        self.items = [_as_bytes(s) for s in flat]
        n = len(self.items)
        # pointer / size tables through numpy: a ctypes array of 2 x 10^4
        # c_char_p costs 15 ms, and a tiled batch repeats a few objects
        addr = {}
        for b in self.items:
            k = id(b)
            if k not in addr:
                addr[k] = ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p).value or 0
        self.ptrs = np.zeros(max(n, 1), dtype=np.uint64)
        self.sizes = np.zeros(max(n, 1), dtype=np.uint64)
        if n:
            self.ptrs[:] = np.fromiter((addr[id(b)] for b in self.items),
                                       dtype=np.uint64, count=n)
            self.sizes[:] = np.fromiter(map(len, self.items), dtype=np.uint64, count=n)
        if tile > 1 and n:
            self.ptrs = np.tile(self.ptrs, tile)
            self.sizes = np.tile(self.sizes, tile)
        self.c = _Strings(self.ptrs.ctypes.data_as(ctypes.POINTER(ctypes.c_char_p)),
                          self.sizes.ctypes.data_as(ctypes.POINTER(ctypes.c_size_t)))


def _flatten2(x):
    if isinstance(x, np.ndarray):
        return list(x.reshape(-1)), (x.shape[0], x.shape[1] if x.ndim > 1 else 0)
    rows = len(x)
    cols = len(x[0]) if rows else 0
    flat = []
    for r in x:
        if len(r) != cols:
            raise InvalidArgumentError("pauli_sums must be rank 2 (ragged rows).")
        flat.extend(r)
    return flat, (rows, cols)


class _Inputs:
    def __init__(self, programs, symbol_names, symbol_values):
        if _rank(programs) != 1:
            raise InvalidArgumentError(
                "programs must be rank 1. Got rank %d." % _rank(programs))
        if _rank(symbol_names) != 1:
            raise InvalidArgumentError(
                "symbol_names must be rank 1. Got rank %d." % _rank(symbol_names))
        # tfq_simulate_ops.py:44: tf.cast(symbol_values, tf.float32)
        vals = np.asarray(symbol_values, dtype=np.float32)
        if vals.ndim != 2:
            raise InvalidArgumentError(
                "symbol_values must be rank 2. Got rank %d." % vals.ndim)
        self.programs = _StringPack(list(programs))
        self.names = _StringPack(list(symbol_names))
        if vals.shape[1] != len(self.names.items):
            raise InvalidArgumentError(
                "Input symbol names and value sizes do not match.")
        self.vals = np.ascontiguousarray(vals)
        self.c = _CircuitInputs(
            self.programs.c, len(self.programs.items), self.names.c,
            len(self.names.items),
            self.vals.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
            self.vals.shape[0])
        self.batch = len(self.programs.items)


def _pauli_pack(pauli_sums):
    if _rank(pauli_sums) != 2:
        raise InvalidArgumentError(
            "pauli_sums must be rank 2. Got rank %d." % _rank(pauli_sums))
    if (isinstance(pauli_sums, list) and len(pauli_sums) > 1 and
            isinstance(pauli_sums[0], (list, tuple)) and
            all(r is pauli_sums[0] for r in pauli_sums)):
        # [sums] * batch: pack one row, repeat the pointer table
        return (_StringPack(list(pauli_sums[0]), tile=len(pauli_sums)),
                len(pauli_sums), len(pauli_sums[0]))
    flat, (rows, cols) = _flatten2(pauli_sums)
    return _StringPack(flat), rows, cols


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


# --------------------------------------------------------------------------
# the five ops
# --------------------------------------------------------------------------
def tfq_simulate_expectation(programs, symbol_names, symbol_values, pauli_sums,
                             *, device: Optional[int] = None) -> np.ndarray:
    """TfqSimulateExpectation (tfq_simulate_ops.py:23-45): float32
    [batch, n_ops]; rows whose program is empty are -2."""
    ctx = get_context(device)
    inp = _Inputs(programs, symbol_names, symbol_values)
    sums, rows, cols = _pauli_pack(pauli_sums)
    out = np.zeros((inp.batch, cols), dtype=np.float32)
    _check(load_library().tfqb_simulate_expectation(
        ctx.handle, ctypes.byref(inp.c), sums.c, rows, cols, _fp(out)))
    return out


def tfq_simulate_state(programs, symbol_names, symbol_values, *,
                       device: Optional[int] = None) -> np.ndarray:
    """TfqSimulateState (tfq_simulate_ops.py:48-68): complex64
    [batch, 2^max_qubits], shorter rows padded with -2."""
    ctx = get_context(device)
    lib = load_library()
    inp = _Inputs(programs, symbol_names, symbol_values)
    job, nmax = ctypes.c_void_p(), ctypes.c_int()
    _check(lib.tfqb_simulate_state_prepare(ctx.handle, ctypes.byref(inp.c),
                                           ctypes.byref(job), ctypes.byref(nmax)))
    try:
        out = np.zeros((inp.batch, 2 ** nmax.value), dtype=np.complex64)
        _check(lib.tfqb_simulate_state_run(job, _fp(out.view(np.float32))))
    finally:
        lib.tfqb_job_free(job)
    return out


def tfq_simulate_samples(programs, symbol_names, symbol_values, num_samples, *,
                         seed: Optional[int] = None, uniforms=None,
                         device: Optional[int] = None) -> np.ndarray:
    """TfqSimulateSamples (tfq_simulate_ops.py:71-99): int8
    [batch, num_samples, max_qubits], -2 padded on the left."""
    ctx = get_context(device)
    lib = load_library()
    inp = _Inputs(programs, symbol_names, symbol_values)
    ns = np.asarray(num_samples)
    if ns.ndim != 1:
        raise InvalidArgumentError(
            "num_samples must be rank 1. Got rank %d." % ns.ndim)
    if ns.shape[0] != 1:
        raise InvalidArgumentError(
            "num_samples must contain 1 element. Got %d." % ns.shape[0])
    S = int(ns[0])
    job, nmax = ctypes.c_void_p(), ctypes.c_int()
    _check(lib.tfqb_simulate_samples_prepare(
        ctx.handle, ctypes.byref(inp.c), S, ctypes.byref(job), ctypes.byref(nmax)))
    try:
        out = np.zeros((inp.batch, S, nmax.value), dtype=np.int8)
        up = None
        if uniforms is not None:
            u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64))
            if u.shape != (inp.batch, S):
                raise InvalidArgumentError("uniforms must be [batch, num_samples]")
            up = u.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        _check(lib.tfqb_simulate_samples_run(
            job, ctypes.c_uint64(seed), up,
            out.ctypes.data_as(ctypes.POINTER(ctypes.c_int8))))
    finally:
        lib.tfqb_job_free(job)
    return out


def tfq_simulate_sampled_expectation(programs, symbol_names, symbol_values,
                                     pauli_sums, num_samples, *,
                                     seed: Optional[int] = None, uniforms=None,
                                     device: Optional[int] = None) -> np.ndarray:
    """TfqSimulateSampledExpectation (tfq_simulate_ops.py:102-135)."""
    ctx = get_context(device)
    inp = _Inputs(programs, symbol_names, symbol_values)
    sums, rows, cols = _pauli_pack(pauli_sums)
    ns = np.asarray(num_samples)
    if ns.ndim != 2:
        raise InvalidArgumentError(
            "num_samples must be rank 2. Got rank %d." % ns.ndim)
    ns = np.ascontiguousarray(ns.astype(np.int32))
    up, ut, us = None, 0, 0
    if uniforms is not None:
        u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64))
        if u.ndim != 4 or u.shape[0] != inp.batch or u.shape[1] != cols:
            raise InvalidArgumentError(
                "uniforms must be [batch, n_ops, terms, shots]")
        up, ut, us = u.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), \
            u.shape[2], u.shape[3]
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    out = np.zeros((inp.batch, cols), dtype=np.float32)
    _check(load_library().tfqb_simulate_sampled_expectation(
        ctx.handle, ctypes.byref(inp.c), sums.c, rows, cols,
        ns.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ns.shape[0],
        ns.shape[1], ctypes.c_uint64(seed), up, ut, us, _fp(out)))
    return out


def tfq_adj_grad(programs, symbol_names, symbol_values, pauli_sums,
                 downstream_grads, *, device: Optional[int] = None) -> np.ndarray:
    """TfqAdjointGradient (tfq_adj_grad_op.py:22-48): float32
    [batch, n_symbols]."""
    ctx = get_context(device)
    inp = _Inputs(programs, symbol_names, symbol_values)
    sums, rows, cols = _pauli_pack(pauli_sums)
    down = np.asarray(downstream_grads, dtype=np.float32)
    if down.ndim != 2:
        raise InvalidArgumentError(
            "downstream_grads must be rank 2. Got rank %d." % down.ndim)
    down = np.ascontiguousarray(down)
    out = np.zeros((inp.batch, len(inp.names.items)), dtype=np.float32)
    _check(load_library().tfqb_adjoint_gradient(
        ctx.handle, ctypes.byref(inp.c), sums.c, rows, cols, _fp(down),
        down.shape[0], down.shape[1], _fp(out)))
    return out


def tfq_calculate_unitary(programs, symbol_names, symbol_values, *,
                          device=None) -> np.ndarray:
    """TfqCalculateUnitary (core/ops/tfq_unitary_op.py:21-53): complex64
    [batch, 2^max_qubits, 2^max_qubits], smaller circuits padded with -2."""
    ctx = get_context(device)
    lib = load_library()
    inp = _Inputs(programs, symbol_names, symbol_values)
    job, nmax = ctypes.c_void_p(), ctypes.c_int()
    _check(lib.tfqb_calculate_unitary_prepare(ctx.handle, ctypes.byref(inp.c),
                                              ctypes.byref(job), ctypes.byref(nmax)))
    try:
        d = 2 ** nmax.value
        out = np.zeros((inp.batch, d, d), dtype=np.complex64)
        _check(lib.tfqb_calculate_unitary_run(job, _fp(out.view(np.float32))))
    finally:
        lib.tfqb_job_free(job)
    return out


def _noisy_expectation(fn_name, programs, symbol_names, symbol_values, pauli_sums,
                       num_samples, seed, uniforms, device):
    ctx = get_context(device)
    inp = _Inputs(programs, symbol_names, symbol_values)
    sums, rows, cols = _pauli_pack(pauli_sums)
    ns = np.asarray(num_samples)
    if ns.ndim != 2:
        raise InvalidArgumentError(
            "num_samples must be rank 2. Got rank %d." % ns.ndim)
    ns = np.ascontiguousarray(ns.astype(np.int32))
    up, ut, uc = None, 0, 0
    if uniforms is not None:
        u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float32))
        if u.ndim != 3 or u.shape[0] != inp.batch:
            raise InvalidArgumentError("uniforms must be [batch, trajectories, channels]")
        up, ut, uc = _fp(u), u.shape[1], u.shape[2]
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    out = np.zeros((inp.batch, cols), dtype=np.float32)
    _check(getattr(load_library(), fn_name)(
        ctx.handle, ctypes.byref(inp.c), sums.c, rows, cols,
        ns.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ns.shape[0],
        ns.shape[1] if ns.ndim > 1 else 0, ctypes.c_uint64(seed), up, ut, uc, _fp(out)))
    return out


def tfq_noisy_expectation(programs, symbol_names, symbol_values, pauli_sums, num_samples,
                          *, seed: Optional[int] = None, uniforms=None,
                          device=None) -> np.ndarray:
    """TfqNoisyExpectation (core/ops/noise/noisy_expectation_op.py:22-70):
    float32 [batch, n_ops], the mean over num_samples[i][j] quantum
    trajectories of the exact expectation value."""
    return _noisy_expectation("tfqb_noisy_expectation", programs, symbol_names,
                              symbol_values, pauli_sums, num_samples, seed, uniforms, device)


def tfq_noisy_sampled_expectation(programs, symbol_names, symbol_values, pauli_sums,
                                  num_samples, *, seed: Optional[int] = None,
                                  uniforms=None, device=None) -> np.ndarray:
    """TfqNoisySampledExpectation (core/ops/noise/noisy_sampled_expectation_op.py):
    every trajectory contributes one measured shot per Pauli term."""
    return _noisy_expectation("tfqb_noisy_sampled_expectation", programs, symbol_names,
                              symbol_values, pauli_sums, num_samples, seed, uniforms, device)


def tfq_noisy_samples(programs, symbol_names, symbol_values, num_samples, *,
                      seed: Optional[int] = None, uniforms=None, measure_uniforms=None,
                      device=None) -> np.ndarray:
    """TfqNoisySamples (core/ops/noise/noisy_samples_op.py): int8
    [batch, num_samples, max_qubits]; every shot is its own trajectory."""
    ctx = get_context(device)
    lib = load_library()
    inp = _Inputs(programs, symbol_names, symbol_values)
    ns = np.asarray(num_samples)
    if ns.ndim != 1:
        raise InvalidArgumentError(
            "num_samples must be rank 1. Got rank %d." % ns.ndim)
    if ns.shape[0] != 1:
        raise InvalidArgumentError(
            "num_samples must contain 1 element. Got %d." % ns.shape[0])
    S = int(ns[0])
    job, nmax = ctypes.c_void_p(), ctypes.c_int()
    _check(lib.tfqb_noisy_samples_prepare(
        ctx.handle, ctypes.byref(inp.c), S, ctypes.byref(job), ctypes.byref(nmax)))
    try:
        out = np.zeros((inp.batch, S, nmax.value), dtype=np.int8)
        up, uc, mp = None, 0, None
        if uniforms is not None:
            u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float32))
            if u.ndim != 3 or u.shape[:2] != (inp.batch, S):
                raise InvalidArgumentError("uniforms must be [batch, num_samples, channels]")
            up, uc = _fp(u), u.shape[2]
        if measure_uniforms is not None:
            mu = np.ascontiguousarray(np.asarray(measure_uniforms, dtype=np.float64))
            if mu.shape != (inp.batch, S):
                raise InvalidArgumentError("measure_uniforms must be [batch, num_samples]")
            mp = mu.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        _check(lib.tfqb_noisy_samples_run(
            job, ctypes.c_uint64(seed), up, uc, mp,
            out.ctypes.data_as(ctypes.POINTER(ctypes.c_int8))))
    finally:
        lib.tfqb_job_free(job)
    return out


def tfq_inner_product(programs, symbol_names, symbol_values, other_programs, *,
                      device: Optional[int] = None) -> np.ndarray:
    """TfqInnerProduct (tensorflow_quantum/core/ops/math_ops/
    inner_product_op.py:72-130): complex64 [batch, n_other],
    out[i][j] = <psi_i(symbol_values[i]) | psi_other[i][j]>."""
    ctx = get_context(device)
    inp = _Inputs(programs, symbol_names, symbol_values)
    if _rank(other_programs) != 2:
        raise InvalidArgumentError(
            "other_programs must be rank 2. Got %d" % _rank(other_programs))
    flat, (rows, cols) = _flatten2(other_programs)
    others = _StringPack(flat)
    out = np.zeros((inp.batch, cols), dtype=np.complex64)
    _check(load_library().tfqb_inner_product(
        ctx.handle, ctypes.byref(inp.c), others.c, rows, cols,
        _fp(out.view(np.float32))))
    return out


def tfq_inner_product_grad(programs, symbol_names, symbol_values, other_programs,
                           prev_grad, *, device: Optional[int] = None) -> np.ndarray:
    """TfqInnerProductGrad as the op returns it (math_ops/
    tfq_inner_product_grad.cc:46-501; inner_product_op.py:23-70 conjugates the
    result for TF's gradient convention): complex64 [batch, n_symbols]."""
    ctx = get_context(device)
    inp = _Inputs(programs, symbol_names, symbol_values)
    if _rank(other_programs) != 2:
        raise InvalidArgumentError(
            "other_programs must be rank 2. Got %d" % _rank(other_programs))
    flat, (rows, cols) = _flatten2(other_programs)
    others = _StringPack(flat)
    down = np.ascontiguousarray(np.asarray(prev_grad, dtype=np.float32))
    if down.ndim != 2:
        raise InvalidArgumentError("downstream_grads must be rank 2.")
    out = np.zeros((inp.batch, len(inp.names.items)), dtype=np.complex64)
    _check(load_library().tfqb_inner_product_grad(
        ctx.handle, ctypes.byref(inp.c), others.c, rows, cols, _fp(down),
        down.shape[0], down.shape[1], _fp(out.view(np.float32))))
    return out


# --------------------------------------------------------------------------
# device-resident jobs (bench.py's kernel-only leg)
# --------------------------------------------------------------------------
class DeviceJob:
    """Parse / plan / upload once; `run()` enqueues only device work."""

    def __init__(self, kind, programs, symbol_names, symbol_values, pauli_sums,
                 downstream_grads=None, device: Optional[int] = None):
        lib = load_library()
        self.ctx = get_context(device)
        inp = _Inputs(programs, symbol_names, symbol_values)
        sums, rows, cols = _pauli_pack(pauli_sums)
        self._job = ctypes.c_void_p()
        if kind == "expectation":
            _check(lib.tfqb_expectation_prepare(
                self.ctx.handle, ctypes.byref(inp.c), sums.c, rows, cols,
                ctypes.byref(self._job)))
            self.shape = (inp.batch, cols)
        elif kind == "adjoint":
            down = np.ascontiguousarray(
                np.asarray(downstream_grads, dtype=np.float32))
            _check(lib.tfqb_adjoint_prepare(
                self.ctx.handle, ctypes.byref(inp.c), sums.c, rows, cols,
                _fp(down), down.shape[0], down.shape[1],
                ctypes.byref(self._job)))
            self.shape = (inp.batch, len(inp.names.items))
        else:
            raise ValueError(kind)

    def run(self):
        _check(load_library().tfqb_job_run_device(self._job))

    def fetch(self) -> np.ndarray:
        out = np.zeros(self.shape, dtype=np.float32)
        _check(load_library().tfqb_job_fetch(self._job, _fp(out)))
        return out

    def close(self):
        if self._job:
            load_library().tfqb_job_free(self._job)
            self._job = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------
# parameter-shift helper ops (reference core/ops/tfq_ps_util_ops.py:20-23):
# host-side rewrites of serialized programs, no GPU and no context
# --------------------------------------------------------------------------
def _rank1(x, name):
    if _rank(x) != 1:
        raise InvalidArgumentError(f"{name} must be rank 1. Got rank {_rank(x)}.")
    return _StringPack(list(x))


def tfq_ps_decompose(programs) -> np.ndarray:
    """TfqPsDecompose: string[B] -> string[B]; parameterised ISwapPow /
    PhasedXPow / FSim / PhasedISwapPow gates become XX / YY / Z / X / CZ powers
    so that every symbol sits in a two-eigenvalue gate."""
    progs = _rank1(programs, "programs")
    out = _StringList()
    _check(load_library().tfqb_ps_decompose(progs.c, len(progs.items), ctypes.byref(out)))
    return np.array(out.take(), dtype=object)


def tfq_ps_symbol_replace(programs, symbols, replacement_symbols) -> np.ndarray:
    """TfqPsSymbolReplace: string[B], string[S], string[S] -> string[B, S, K];
    entry (i, j, k) is programs[i] with the kth occurrence of symbols[j]
    renamed to replacement_symbols[j]; padded with empty programs."""
    progs = _rank1(programs, "programs")
    syms = _rank1(symbols, "symbols")
    reps = _rank1(replacement_symbols, "replacement_symbols")
    out = _StringList()
    pad = ctypes.c_int(0)
    _check(load_library().tfqb_ps_symbol_replace(
        progs.c, len(progs.items), syms.c, len(syms.items), reps.c, len(reps.items),
        ctypes.byref(out), ctypes.byref(pad)))
    flat = out.take()
    res = np.empty((len(progs.items), len(syms.items), pad.value), dtype=object)
    res.reshape(-1)[:] = flat
    return res


def tfq_ps_weights_from_symbols(programs, symbols) -> np.ndarray:
    """TfqPsWeightsFromSymbols: string[B], string[S] -> float32[B, S, K]: the
    exponent_scalar of every operation whose exponent is that symbol."""
    progs = _rank1(programs, "programs")
    syms = _rank1(symbols, "symbols")
    w = ctypes.POINTER(ctypes.c_float)()
    pad = ctypes.c_int(0)
    lib = load_library()
    _check(lib.tfqb_ps_weights_from_symbols(progs.c, len(progs.items), syms.c, len(syms.items),
                                            ctypes.byref(w), ctypes.byref(pad)))
    shape = (len(progs.items), len(syms.items), pad.value)
    n = int(np.prod(shape))
    res = (np.ctypeslib.as_array(w, shape=(max(n, 1),))[:n].reshape(shape).copy()
           if n else np.zeros(shape, np.float32))
    lib.tfqb_free_floats(w)
    return res


# --------------------------------------------------------------------------
# host-only helpers (no GPU needed)
# --------------------------------------------------------------------------
def host_gate_matrix(kind: int, params, grad_param: int = -1) -> np.ndarray:
    p = np.ascontiguousarray(np.asarray(params, dtype=np.float32))
    out = np.zeros(32, dtype=np.float32)
    _check(load_library().tfqb_host_gate_matrix(kind, _fp(p), len(p),
                                                grad_param, _fp(out)))
    two = kind == 1 or 6 <= kind <= 12 or kind in (14, 15)
    dim = 4 if two else 2
    return out[:2 * dim * dim].view(np.complex64).reshape(dim, dim).copy()


def host_describe_plan(program, symbol_names=(), adjoint=False) -> dict:
    lib = load_library()
    prog = _as_bytes(program)
    names = _StringPack(list(symbol_names))
    out = ctypes.c_char_p()
    _check(lib.tfqb_host_describe_plan(prog, len(prog), names.c,
                                       len(names.items), 1 if adjoint else 0,
                                       ctypes.byref(out)))
    try:
        return json.loads(out.value.decode())
    finally:
        lib.tfqb_free_string(ctypes.cast(out, ctypes.c_void_p))


def host_jit_source(program, symbol_names=(), adjoint=False, pass_index=0,
                    phase_free=False) -> str:
    """CUDA C++ text of the run-time specialised kernel of one pass (csrc/jit.h);
    '' when the pass is not specialisable."""
    lib = load_library()
    prog = _as_bytes(program)
    names = _StringPack(list(symbol_names))
    out = ctypes.c_char_p()
    _check(lib.tfqb_host_jit_source(prog, len(prog), names.c, len(names.items),
                                    1 if adjoint else (2 if phase_free else 0), pass_index,
                                    ctypes.byref(out)))
    try:
        return out.value.decode()
    finally:
        lib.tfqb_free_string(ctypes.cast(out, ctypes.c_void_p))


def host_jit_expect_source(program, pauli_sums, pass_index=0) -> str:
    """CUDA C++ text of the specialised kernel of one tile pass of the
    PauliSum expectation plan (csrc/jit.h); '' when not specialisable."""
    lib = load_library()
    prog = _as_bytes(program)
    sums = _StringPack([_as_bytes(p) for p in pauli_sums])
    out = ctypes.c_char_p()
    _check(lib.tfqb_host_jit_expect_source(prog, len(prog), sums.c,
                                           len(sums.items), pass_index,
                                           ctypes.byref(out)))
    try:
        return out.value.decode()
    finally:
        lib.tfqb_free_string(ctypes.cast(out, ctypes.c_void_p))


def host_describe_pauli_sum(program, pauli_sum) -> dict:
    lib = load_library()
    prog, ps = _as_bytes(program), _as_bytes(pauli_sum)
    out = ctypes.c_char_p()
    _check(lib.tfqb_host_describe_pauli_sum(prog, len(prog), ps, len(ps),
                                            ctypes.byref(out)))
    try:
        return json.loads(out.value.decode())
    finally:
        lib.tfqb_free_string(ctypes.cast(out, ctypes.c_void_p))


def host_describe_sharded(program, symbol_names, pauli_sums, world: int) -> dict:
    lib = load_library()
    prog = _as_bytes(program)
    names = _StringPack(list(symbol_names))
    sums = _StringPack(list(pauli_sums))
    out = ctypes.c_char_p()
    _check(lib.tfqb_host_describe_sharded(
        prog, len(prog), names.c, len(names.items), sums.c, len(sums.items),
        world, ctypes.byref(out)))
    try:
        return json.loads(out.value.decode())
    finally:
        lib.tfqb_free_string(ctypes.cast(out, ctypes.c_void_p))
