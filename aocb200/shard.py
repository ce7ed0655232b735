"""Sequence-level sharding: sequences are independent (networks/engine/eval_manager_mm.py:172, state reset per
sequence :182-193), frames inside one are not.  Sequence i goes to rank i mod world_size; there is no collective on
the per-frame path -- only an optional one-time weight broadcast at init (rank 0 -> all).

Bank sharding is the one exception, for ONE long sequence on several GPUs: every rank runs the same
frames, the global matching against the memory bank (matching.py:2384-2516) is split by bank row blocks, and the
partial minima are exchanged from inside the matching kernel through peer-mapped memory (`setup_bank_sharding`)."""
import torch


def sequences_for_rank(num_sequences, rank, world_size):
    return list(range(rank, num_sequences, world_size))


def broadcast_state_dict(state_dict, src=0, device=None):
    """One flat buffer, one collective (NCCL over NVLink on GPUs, gloo in CPU tests).  In place; returns state_dict."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return state_dict
    keys = list(state_dict.keys())
    dev = device if device is not None else state_dict[keys[0]].device
    flat = torch.cat([state_dict[k].detach().reshape(-1).to(device=dev, dtype=torch.float32) for k in keys])
    dist.broadcast(flat, src=src)
    off = 0
    for k in keys:
        n = state_dict[k].numel()
        state_dict[k].copy_(flat[off:off + n].view_as(state_dict[k]).to(state_dict[k].device))
        off += n
    return state_dict


def gather_results(obj, dst=0):
    """Collect one small python object per rank on rank `dst` (end-of-run statistics, never per frame)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out


def shard_row_blocks(n_row_blocks, rank, world_size):
    """[begin, end) of the bank row blocks (256 sorted rows each) rank `rank` contracts: the same split as
    aoc_match_shard_range -- contiguous, disjoint, covering, sizes differing by at most one block."""
    return n_row_blocks * rank // world_size, n_row_blocks * (rank + 1) // world_size


def exchange_handles(handle):
    """all-gather of one small bytes object per rank (the 64-byte CUDA IPC handle of the rank's exchange area)"""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, bytes(handle))
    return out


def setup_bank_sharding(engine, cap_hw):
    """Gives `engine` (one per rank, all loaded with the same weights) the peer-mapped exchange areas of the
    bank-sharded global matching: allocates this rank's area, exchanges the IPC handles over torch.distributed (init
    time only), maps the peers' areas.  Afterwards engine.forward_for_eval contracts only this rank's share of the bank
    and the kernels exchange the partial minima over NVLink; every rank must be fed the same frames, label maps and
    numpy seed.  cap_hw: largest h*w (stride-4 feature pixels) that will be matched."""
    import ctypes
    import torch
    import torch.distributed as dist
    L = engine.L
    rank, world = dist.get_rank(), dist.get_world_size()
    assert 2 <= world <= 9, "bank sharding supports 2..9 ranks"
    nbytes = L.match_shard_area_bytes(world, int(cap_hw))
    own = ctypes.c_void_p()
    L.peer_alloc(nbytes, ctypes.byref(own))
    handle = ctypes.create_string_buffer(64)
    L.peer_export(own, handle)
    handles = exchange_handles(handle.raw)
    areas = (ctypes.c_void_p * world)()
    for g in range(world):
        if g == rank:
            areas[g] = own.value
        else:
            p = ctypes.c_void_p()
            L.peer_open(handles[g], ctypes.byref(p))
            areas[g] = p.value
    dist.barrier()                                   # every area is mapped everywhere before the first frame
    engine.shard = dict(rank=rank, world=world, areas=areas, cap_hw=int(cap_hw), own=own,
                        state=torch.zeros(16, dtype=torch.int32, device=engine.dev))
    engine._static.clear()                           # graphs captured with the unsharded matching
    return engine.shard
