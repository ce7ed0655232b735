"""Sequence-level sharding: sequences are independent (networks/engine/eval_manager_mm.py:172, state reset per
sequence :182-193), frames inside one are not.  Sequence i goes to rank i mod world_size; there is no collective on
the per-frame path -- only an optional one-time weight broadcast at init (rank 0 -> all)."""
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
