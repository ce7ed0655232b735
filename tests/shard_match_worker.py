"""Worker of tests/test_gpu_shard_match.py (one process per GPU, launched with torch.distributed.run): the same clip
through the engine (a) unsharded and (b) with the bank sharded over the ranks; the logits must be bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from aocb200.model import get_module
    from aocb200.params import synthetic_state_dict
    from aocb200.sequence import run_sequence
    from aocb200.shard import setup_bank_sharding
    from aocb200.synth import make_clip
    rank, world = dist.get_rank(), dist.get_world_size()
    K, H, W, T = 3, 129, 225, 7
    frames, labels = make_clip(5, H, W, K, T)
    model = get_module()(None, None)
    model.load_state_dict(synthetic_state_dict(1234))
    model = model.cuda(local).eval()
    eng = model.engine()

    def run(graphs):
        eng.use_graphs = graphs
        logits = []
        np.random.seed(17)
        preds = run_sequence(model, frames, labels[0], K, mem_every=2, device=dev,
                             on_frame=lambda t, p, y: logits.append(eng.last_logits.clone()))
        torch.cuda.synchronize()
        return preds, logits

    want_p, want_l = run(True)
    h, w = (H + 3) // 4, (W + 3) // 4
    setup_bank_sharding(eng, cap_hw=h * w)
    ok = True
    for graphs in (True, False):
        got_p, got_l = run(graphs)
        for t, (a, b) in enumerate(zip(got_l, want_l)):
            same = torch.equal(a, b) and torch.equal(got_p[t], want_p[t])
            ok = ok and same
            if rank == 0:
                print("[parity] bank sharded over %d GPUs (%s) frame %d, bank %d frames: logits bit-identical to the "
                      "single-GPU run: %s" % (world, "graphs" if graphs else "plain launches", t + 1, 1 + t // 2, same))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
