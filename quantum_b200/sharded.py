"""One state vector sharded over world = 2^g ranks by its top ("global")
qubits, with all-to-all global<->local qubit swaps.

Each rank holds 2^(n-g) amplitudes; the rank index supplies the top g
amplitude-index bits.  Gates on local bits need no communication; diagonal
gates / controls / Z-type Pauli factors on rank bits are signs and predicates
computed from the rank.  When a dense gate (or an X/Y Pauli factor) needs a
qubit that is currently a rank bit, the planner (csrc/plan.cc PlanSharded)
inserts ONE exchange: the g rank bits trade places with the top g local bits,
which is exactly `all_to_all_single` with equal splits over the flat shard.

Drivers of the same stage list:
  * `peer_sharded_expectation` : one process per GPU; the exchange happens
    INSIDE the library through peer memory (CUDA IPC over NVLink / NVSwitch:
    each rank's kernels load their incoming chunks from the peers' HBM,
    ordered by epoch flags in device memory).  torch.distributed only carries
    the 256-byte handles once per job.  This is the product path.
  * `sharded_expectation`  : the same stages with the exchange done by the
    host as `all_to_all_single` (NCCL) — kept as the library baseline the
    peer-memory exchange is measured against;
  * `emulated_peer_sharded_expectation` / `emulated_sharded_expectation` :
    all virtual ranks in one process on one GPU (one stream per rank) — the
    tests' way to run the planner, the kernels and the flag protocol without
    a multi-GPU box.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import numpy as np

from . import ops


class _DevBuf:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch
    can wrap it without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False),
            "version": 3, "strides": None}


class ShardedJob:
    def __init__(self, program, symbol_names, symbol_values, pauli_sums,
                 world: int, rank: int, device: Optional[int] = None,
                 ctx: Optional["ops.Context"] = None):
        lib = ops.load_library()
        self.ctx = ctx if ctx is not None else ops.get_context(device)
        vals = np.asarray(symbol_values, dtype=np.float32).reshape(1, -1)
        inp = ops._Inputs([program], symbol_names, vals)
        sums = ops._StringPack(list(pauli_sums))
        self.n_ops = len(sums.items)
        self._job = ctypes.c_void_p()
        ns, nt = ctypes.c_int(), ctypes.c_int()
        ops._check(lib.tfqb_sharded_prepare(
            self.ctx.handle, ctypes.byref(inp.c), sums.c, self.n_ops, world,
            rank, ctypes.byref(self._job), ctypes.byref(ns), ctypes.byref(nt)))
        self.n_stages, self.n_terms = ns.value, nt.value
        self.kinds = [lib.tfqb_sharded_stage_kind(self._job, i)
                      for i in range(self.n_stages)]
        self.world, self.rank = world, rank

    def run_stage(self, i: int):
        ops._check(ops.load_library().tfqb_sharded_run_stage(self._job, i))

    def buffers(self):
        send, recv = ctypes.c_void_p(), ctypes.c_void_p()
        nbytes = ctypes.c_size_t()
        ops._check(ops.load_library().tfqb_sharded_buffers(
            self._job, ctypes.byref(send), ctypes.byref(recv),
            ctypes.byref(nbytes)))
        return send.value, recv.value, nbytes.value

    def tensors(self):
        """(send, recv) float32 torch views of the two shard buffers."""
        import torch
        send, recv, nbytes = self.buffers()
        dev = "cuda:%d" % self.ctx.device
        return (torch.as_tensor(_DevBuf(send, nbytes), device=dev),
                torch.as_tensor(_DevBuf(recv, nbytes), device=dev))

    def partials(self) -> np.ndarray:
        out = np.zeros(max(self.n_terms, 1), dtype=np.float64)
        ops._check(ops.load_library().tfqb_sharded_partials(
            self._job, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        return out[:self.n_terms]

    def finish(self, totals: np.ndarray) -> np.ndarray:
        t = np.ascontiguousarray(np.asarray(totals, dtype=np.float64))
        if t.size == 0:
            t = np.zeros(1)
        out = np.zeros(self.n_ops, dtype=np.float32)
        ops._check(ops.load_library().tfqb_sharded_finish(
            self._job, t.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
        return out

    # ---- exchange inside the library (peer memory) ----------------------
    def export(self) -> bytes:
        buf = ctypes.create_string_buffer(ops.PEER_HANDLE_BYTES)
        ops._check(ops.load_library().tfqb_sharded_export(self._job, buf))
        return buf.raw

    def connect(self, handles: List[bytes]):
        blob = b"".join(handles)
        assert len(blob) == self.world * ops.PEER_HANDLE_BYTES
        ops._check(ops.load_library().tfqb_sharded_connect(self._job, blob, self.world))

    def enqueue(self):
        ops._check(ops.load_library().tfqb_sharded_enqueue(self._job))

    def result(self) -> np.ndarray:
        out = np.zeros(self.n_ops, dtype=np.float32)
        ops._check(ops.load_library().tfqb_sharded_result(
            self._job, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
        return out

    def sample(self, num_samples: int, n_qubits: int, seed: int = 0, uniforms=None):
        """This rank's shots of `num_samples` draws from the sharded state
        (after enqueue): (int8 [num_samples, n_qubits], owned mask)."""
        out = np.zeros((num_samples, n_qubits), dtype=np.int8)
        own = np.zeros(num_samples, dtype=np.int32)
        up = None
        if uniforms is not None:
            u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64))
            up = u.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        ops._check(ops.load_library().tfqb_sharded_sample(
            self._job, num_samples, ctypes.c_uint64(seed), up,
            out.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
            own.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        return out, own

    def stats(self) -> dict:
        st = ops.ExchangeStats()
        ops._check(ops.load_library().tfqb_sharded_stats(self._job, ctypes.byref(st)))
        return st.as_dict()

    def close(self):
        if self._job:
            ops.load_library().tfqb_job_free(self._job)
            self._job = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sharded_expectation(program, symbol_names, symbol_values, pauli_sums,
                        group=None, device: Optional[int] = None,
                        stats: Optional[dict] = None) -> np.ndarray:
    """<psi| O_j |psi> for ONE program whose state is sharded over the ranks
    of `group` (torch.distributed, NCCL).  Every rank passes identical inputs
    and gets the identical float32[n_ops] result."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    job = ShardedJob(program, symbol_names, symbol_values, pauli_sums, world,
                     rank, device)
    dev = torch.device("cuda", job.ctx.device)
    n_exch, exch_s = 0, 0.0
    try:
        for i, kind in enumerate(job.kinds):
            if kind == 1:
                job.ctx.sync()
                send, recv = job.tensors()
                ev0 = torch.cuda.Event(enable_timing=True)
                ev1 = torch.cuda.Event(enable_timing=True)
                ev0.record()
                dist.all_to_all_single(recv, send, group=group)
                ev1.record()
                torch.cuda.synchronize(dev)
                exch_s += ev0.elapsed_time(ev1) * 1e-3
                n_exch += 1
            job.run_stage(i)
        part = torch.from_numpy(job.partials().copy()).to(dev)
        if part.numel():
            dist.all_reduce(part, group=group)
        out = job.finish(part.cpu().numpy())
        if stats is not None:
            _, _, nbytes = job.buffers()
            stats.update(exchanges=n_exch, exchange_seconds=exch_s,
                         shard_bytes=nbytes, stages=list(job.kinds))
        return out
    finally:
        job.close()


def emulated_sharded_expectation(program, symbol_names, symbol_values,
                                 pauli_sums, world: int,
                                 device: Optional[int] = None,
                                 stats: Optional[dict] = None) -> np.ndarray:
    """All `world` virtual ranks on one GPU; the exchange copies chunk c of
    rank r's shard into chunk r of rank c's alternate buffer."""
    jobs: List[ShardedJob] = [
        ShardedJob(program, symbol_names, symbol_values, pauli_sums, world, r,
                   device) for r in range(world)]
    try:
        for i, kind in enumerate(jobs[0].kinds):
            if kind == 1:
                jobs[0].ctx.sync()
                views = [j.tensors() for j in jobs]
                chunk = views[0][0].numel() // world
                for r in range(world):
                    for c in range(world):
                        views[c][1][r * chunk:(r + 1) * chunk].copy_(
                            views[r][0][c * chunk:(c + 1) * chunk])
                import torch
                torch.cuda.synchronize()
            for j in jobs:
                j.run_stage(i)
        total = np.sum([j.partials() for j in jobs], axis=0) \
            if jobs[0].n_terms else np.zeros(0)
        if stats is not None:
            stats.update(stages=list(jobs[0].kinds),
                         exchanges=sum(1 for k in jobs[0].kinds if k == 1))
        return jobs[0].finish(total)
    finally:
        for j in jobs:
            j.close()


def peer_sharded_job(program, symbol_names, symbol_values, pauli_sums, group=None,
                     device: Optional[int] = None) -> ShardedJob:
    """Prepare + export + all-gather the handles + connect: a job whose
    `enqueue()` / `result()` run every stage, exchanges included, inside the
    library.  Collective: every rank of `group` calls it with identical
    inputs."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    job = ShardedJob(program, symbol_names, symbol_values, pauli_sums, world, rank, device)
    mine = torch.frombuffer(bytearray(job.export()), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        mine = mine.to(torch.device("cuda", job.ctx.device))
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    job.connect([bytes(p.cpu().numpy().tobytes()) for p in parts])
    return job


def peer_sharded_expectation(program, symbol_names, symbol_values, pauli_sums,
                             group=None, device: Optional[int] = None,
                             stats: Optional[dict] = None) -> np.ndarray:
    """<psi| O_j |psi> for ONE program whose state is sharded over the ranks
    of `group`, the qubit swaps done through peer memory inside the library.
    Every rank passes identical inputs and gets the identical float32[n_ops]."""
    job = peer_sharded_job(program, symbol_names, symbol_values, pauli_sums, group, device)
    try:
        job.enqueue()
        out = job.result()
        if stats is not None:
            stats.update(job.stats(), stages=list(job.kinds))
        return out
    finally:
        # nobody may unmap a shard a peer is still reading
        import torch.distributed as dist
        dist.barrier(group)
        job.close()


def emulated_peer_sharded_expectation(program, symbol_names, symbol_values, pauli_sums,
                                      world: int, device: Optional[int] = None,
                                      repeats: int = 1,
                                      stats: Optional[dict] = None) -> List[np.ndarray]:
    """All `world` ranks in this process on one GPU, one context (stream) per
    rank: the same library path as `peer_sharded_expectation` — flags, waits
    and pulls included — with raw pointers in place of IPC mappings.  Returns
    every rank's result of every repeat (they must all be identical).

    Ranks that share a process need one hardware queue per rank's stream
    (`CUDA_DEVICE_MAX_CONNECTIONS` >= world + a few, set before CUDA
    initialises; default 8): in a shared queue one rank's kernels can be stuck
    behind another rank's spinning wait kernel."""
    dev = ops.default_device() if device is None else device
    ctxs = [ops.Context(dev) for _ in range(world)]
    jobs: List[ShardedJob] = []
    try:
        for r in range(world):
            jobs.append(ShardedJob(program, symbol_names, symbol_values, pauli_sums,
                                   world, r, ctx=ctxs[r]))
        handles = [j.export() for j in jobs]
        for j in jobs:
            j.connect(handles)
        outs = []
        for _ in range(repeats):
            for j in jobs:
                j.enqueue()
            outs += [j.result() for j in jobs]
        if stats is not None:
            stats.update(jobs[0].stats(), stages=list(jobs[0].kinds))
        return outs
    finally:
        for j in jobs:
            j.close()
        for c in ctxs:
            c.close()


def peer_sharded_samples(program, symbol_names, symbol_values, num_samples: int,
                         n_qubits: int, seed: int = 0, group=None,
                         device: Optional[int] = None) -> np.ndarray:
    """TfqSimulateSamples for ONE program whose state is sharded over the ranks
    of `group`: int8 [num_samples, n_qubits], identical on every rank."""
    import torch
    import torch.distributed as dist
    job = peer_sharded_job(program, symbol_names, symbol_values, [], group, device)
    try:
        job.enqueue()
        out, own = job.sample(num_samples, n_qubits, seed)
        dev = torch.device("cuda", job.ctx.device)
        t = torch.from_numpy(out.astype(np.int32) * own[:, None]).to(dev)
        c = torch.from_numpy(own.copy()).to(dev)
        dist.all_reduce(t, group=group)
        dist.all_reduce(c, group=group)
        assert int(c.min()) == 1 and int(c.max()) == 1, "every shot belongs to one rank"
        return t.cpu().numpy().astype(np.int8)
    finally:
        dist.barrier(group)
        job.close()


def emulated_peer_sharded_samples(program, symbol_names, symbol_values, num_samples: int,
                                  n_qubits: int, world: int, seed: int = 0, uniforms=None,
                                  device: Optional[int] = None):
    """All ranks in this process on one GPU (see
    emulated_peer_sharded_expectation).  Returns (samples, owner count)."""
    dev = ops.default_device() if device is None else device
    ctxs = [ops.Context(dev) for _ in range(world)]
    jobs: List[ShardedJob] = []
    try:
        for r in range(world):
            jobs.append(ShardedJob(program, symbol_names, symbol_values, [], world, r,
                                   ctx=ctxs[r]))
        handles = [j.export() for j in jobs]
        for j in jobs:
            j.connect(handles)
        for j in jobs:
            j.enqueue()
        for j in jobs:
            j.result()
        # the norm exchange waits for every rank: enqueue all before any host sync
        import threading
        res = [None] * world

        def work(r):
            res[r] = jobs[r].sample(num_samples, n_qubits, seed, uniforms)
        th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        total = np.zeros((num_samples, n_qubits), dtype=np.int32)
        count = np.zeros(num_samples, dtype=np.int32)
        for out, own in res:
            total += out.astype(np.int32) * own[:, None]
            count += own
        return total.astype(np.int8), count
    finally:
        for j in jobs:
            j.close()
        for c in ctxs:
            c.close()
