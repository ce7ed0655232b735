"""Sequence sharding and the one-time weight broadcast (aocb200/shard.py) with world_size 2 on the gloo backend.
Mirrors the only multi-process structure of the path: one process per GPU, independent sequences, no per-frame
collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aocb200.shard import broadcast_state_dict, gather_results, sequences_for_rank


def test_partition_is_disjoint_and_complete():
    for n, w in ((8, 8), (8, 2), (5, 4), (3, 8), (0, 2)):
        parts = [sequences_for_rank(n, r, w) for r in range(w)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)           # every rank starts from different weights
        sd = {"a.weight": torch.randn(7, 3, generator=g), "b.bias": torch.randn(5, generator=g),
              "c.running_var": torch.rand(4, generator=g)}
        broadcast_state_dict(sd, src=0)
        g0 = torch.Generator().manual_seed(100)
        want = {"a.weight": torch.randn(7, 3, generator=g0), "b.bias": torch.randn(5, generator=g0),
                "c.running_var": torch.rand(4, generator=g0)}
        ok = all(torch.equal(sd[k], want[k]) for k in sd)
        mine = sequences_for_rank(5, rank, world)
        res = gather_results({"rank": rank, "seqs": mine, "ok": ok}, dst=0)
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r["rank"] for r in res] == [0, 1]
    assert all(r["ok"] for r in res), "rank-0 weights did not arrive bit-exactly"
    assert sorted(res[0]["seqs"] + res[1]["seqs"]) == [0, 1, 2, 3, 4]


def test_bank_row_block_split_matches_the_library():
    """the host-side split (shard_row_blocks) is the one the kernel launcher uses (aoc_match_shard_range): contiguous,
    disjoint, covering, balanced to one 256-row block"""
    import ctypes
    from aocb200.lib import lib
    from aocb200.shard import shard_row_blocks
    L = lib()
    for nrb in (0, 1, 2, 7, 101, 2021):
        for world in (2, 3, 8):
            parts = [shard_row_blocks(nrb, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == nrb
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
            for r in range(world):
                a, b = ctypes.c_int(), ctypes.c_int()
                L.match_shard_range(nrb * 256, r, world, ctypes.byref(a), ctypes.byref(b))
                assert (a.value, b.value) == parts[r]


def _handle_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aocb200.shard import exchange_handles
        mine = bytes([rank]) * 64                                # stand-in for a 64-byte CUDA IPC handle
        got = exchange_handles(mine)
        if rank == 0:
            q.put(got)
    finally:
        dist.destroy_process_group()


def test_ipc_handle_exchange_world2():
    """the init-time exchange of the exchange areas' IPC handles (the only host collective of the bank-sharded mode)"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_handle_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [bytes([0]) * 64, bytes([1]) * 64]
