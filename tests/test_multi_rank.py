"""world_size-2 gloo test of the batch-sharding host logic (no GPU): every
row is handled by exactly one rank and the gathered result is in row order."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from quantum_b200 import sharding


def test_row_blocks_partition_the_batch():
    for batch in (0, 1, 5, 16, 4097):
        for world in (1, 2, 3, 8):
            blocks = [sharding.row_block(batch, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == batch
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _fake_op(programs, symbol_names, symbol_values, sums):
    # stands in for a device op: one output row per input row
    return np.asarray([[len(p) + float(v.sum()) + len(s[0])]
                       for p, v, s in zip(programs, symbol_values, sums)],
                      dtype=np.float32)


def _fake_state_op(programs, symbol_names, symbol_values):
    nmax = max(len(p) for p in programs)
    out = np.full((len(programs), 2 ** nmax), -2, np.complex64)
    for i, p in enumerate(programs):
        out[i, :2 ** len(p)] = i + 1
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 7
    programs = [b"x" * (i + 1) for i in range(B)]
    vals = np.arange(B * 3, dtype=np.float32).reshape(B, 3)
    sums = [[b"s" * (i % 3)] for i in range(B)]
    full = sharding.run_sharded(_fake_op, programs, ["a", "b", "c"], vals, sums)
    local = sharding.run_sharded(_fake_op, programs, ["a", "b", "c"], vals, sums,
                                 gather=False)
    small = [b"q" * (1 + (i % 3)) for i in range(B)]
    states = sharding.run_sharded(_fake_state_op, small, [], np.zeros((B, 0)),
                                  pad_value=-2)
    if rank == 0:
        q.put((full, local, states))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_call_matches_single_rank():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, local, states = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    B = 7
    programs = [b"x" * (i + 1) for i in range(B)]
    vals = np.arange(B * 3, dtype=np.float32).reshape(B, 3)
    sums = [[b"s" * (i % 3)] for i in range(B)]
    ref = _fake_op(programs, None, vals, sums)
    np.testing.assert_array_equal(full, ref)
    np.testing.assert_array_equal(local, ref[:4])      # rank 0 owns rows 0..3
    assert states.shape == (B, 8)
    # every row keeps its own amplitudes and is padded with -2 on the right
    for i in range(B):
        n = 1 + (i % 3)
        assert (states[i, 2 ** n:] == -2).all() and (states[i, :2 ** n] != -2).all()
