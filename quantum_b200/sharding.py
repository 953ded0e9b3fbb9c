"""Batch sharding of the circuit-execution ops over ranks / GPUs.

Rows of `programs`, `symbol_values`, `pauli_sums`, `num_samples` and
`downstream_grads` are independent — the reference itself parallelises over
them (tfq_simulate_expectation_op.cc:247-248, tfq_adj_grad_op.cc:282-283) —
so each rank simulates one contiguous block of rows on its own GPU and there
is NO data-path collective.  The only communication is the optional gather of
the (tiny) result tensors, done with torch.distributed (NCCL on GPUs, gloo in
the CPU tests).
"""
from typing import Callable, Optional, Sequence, Tuple

import numpy as np


def row_block(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most 1."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def run_sharded(op: Callable, programs: Sequence, symbol_names: Sequence,
                symbol_values, *row_args, rank: Optional[int] = None,
                world: Optional[int] = None, gather: bool = True,
                pad_value=None, context=None, **kw):
    """Call `op(programs, symbol_names, symbol_values, *row_args)` on this
    rank's block of rows; with `gather`, all-gather the per-rank results
    (concatenated in rank order = original row order).

    Ops whose trailing output dims depend on the rows (TfqSimulateState /
    TfqSimulateSamples pad to the batch-wide max qubit count) are re-padded
    to the widest rank with `pad_value` (-2, as the reference pads).

    `context` (an ops.Context): told the global index of this rank's first
    row, so that the sampling ops draw, for a given seed, the uniforms of the
    unsplit batch (their Philox streams are keyed by global row)."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    vals = np.asarray(symbol_values)
    lo, hi = row_block(len(programs), rank, world)
    if context is not None:
        context.set_row_offset(lo)
    local = op(list(programs[lo:hi]), symbol_names, vals[lo:hi],
               *[a[lo:hi] for a in row_args], **kw)
    if not gather or world == 1:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, local)
    if pad_value is not None:
        width = max(p.shape[-1] for p in parts)
        out = []
        for p in parts:
            if p.shape[-1] < width:
                padw = [(0, 0)] * (p.ndim - 1)
                # samples pad on the left, states on the right
                padw.append((width - p.shape[-1], 0) if p.dtype == np.int8
                            else (0, width - p.shape[-1]))
                p = np.pad(p, padw, constant_values=pad_value)
            out.append(p)
        parts = out
    return np.concatenate(parts, axis=0)
