"""Saves the frame-1 logits of a synthetic clip under the current AOCB200_OPTS (numerics A/B of kernel modes)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from aocb200.model import get_module
    from aocb200.params import synthetic_state_dict
    from aocb200.sequence import run_sequence
    from aocb200.synth import make_clip, restrict_size
    out, Hn, Wn, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    H, W = restrict_size(Hn, Wn, 10 ** 9)
    frames, labels = make_clip(11, H, W, K, 2)
    model = get_module()(None, None)
    model.load_state_dict(synthetic_state_dict(1234))
    model = model.cuda(0).eval()
    model.engine().use_graphs = False
    model.engine().keep_debug = True
    np.random.seed(11)
    run_sequence(model, frames, labels[0], K, mem_every=2, device=torch.device("cuda:0"))
    d = model.engine().debug
    torch.save({"logits": model.engine().last_logits.cpu(), "labels": d["labels"].cpu(), "cent": d["cent"].cpu(),
                "g": d["g"].cpu(), "gc": d["gc"].cpu(), "loc": d["loc"].buf.cpu(), "S": d["S"].cpu()[:4000000]}, out)


if __name__ == "__main__":
    main()
