"""Host logic of the per-sequence driver (aocb200/sequence.py::run_sequence, the restatement of
eval_manager_mm.py:172-361) with a stub model: label-existence filter, uncertainty -> label 125 on memory frames,
objects joining with ground truth at a later frame."""
import math

import torch

from aocb200.sequence import run_sequence


class Stub:
    """forward_for_eval that returns fixed probabilities and records what the loop hands back"""

    def __init__(self, probs):
        self.probs, self.calls = probs, []

    def forward_for_eval(self, memory, ref_e, ref_m, prev_e, prev_m, img, pred_size, gt_ids):
        self.calls.append(dict(n_ref=len(ref_e), ref_m=[m.clone() for m in ref_m],
                               prev_m=None if prev_m is None else prev_m.clone()))
        emb = torch.zeros(1, 100, 2, 2)
        return (None if prev_e is None else self.probs.clone()), emb, memory


def test_eval_loop_bookkeeping():
    H = W = 4
    p = torch.zeros(1, 3, H, W)
    p[0, 0], p[0, 1], p[0, 2] = 0.2, 0.3, 0.5                 # id 2 would win everywhere ...
    p[0, :, 0, 0] = torch.tensor([0.98, 0.01, 0.01])          # ... except one confident background pixel
    first = torch.zeros(H, W, dtype=torch.long)
    first[1, 1] = 1                                           # id 2 is absent from the first frame
    stub = Stub(p)
    join = torch.zeros(H, W, dtype=torch.long)
    join[3, 3] = 2
    preds = run_sequence(stub, torch.zeros(5, 3, H, W), first, 2, mem_every=2, unc_ratio=0.5, later_labels={3: join})
    # frames 1, 2: id 2 filtered out (eval_manager_mm.py:252-261) -> id 1 wins
    assert (preds[0][1:, :] == 1).all() and preds[0][0, 0] == 0
    # frame 2 is a memory frame: entropy over the EXISTING ids; -(0.2 ln 0.2 + 0.3 ln 0.3) = 0.683 > 0.5 -> 125
    c = stub.calls[3]["ref_m"][1][0, 0]
    ent = -(0.2 * math.log(0.2 + 1e-6) + 0.3 * math.log(0.3 + 1e-6))
    assert ent > 0.5 and c[1, 1] == 125 and c[0, 0] == 0
    # frame 3 carries ground truth for id 2: joined into the prediction, the confident mask and the previous mask
    assert preds[2][3, 3] == 2 and (preds[2][1, :] == 1).all()
    assert stub.calls[4]["n_ref"] == 3 and stub.calls[4]["ref_m"][2][0, 0, 3, 3] == 2
    assert stub.calls[4]["prev_m"][0, 0, 3, 3] == 2
    # frame 4: id 2 has been seen now -> it wins wherever the stub says so
    assert (preds[3][1:, :] == 2).all() and preds[3][0, 0] == 0
