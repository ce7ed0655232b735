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


class _Replay:
    """the recording stub of tools/make_eval_loop_golden.py: frame t's probabilities are fixed, every call is recorded"""

    def __init__(self, probs):
        self.probs, self.calls, self.t = probs, [], 0

    def forward_for_eval(self, memory, ref_e, ref_m, prev_e, prev_m, img, pred_size=None, gt_ids=None):
        self.calls.append(dict(n_ref=len(ref_e), ref_m=[m.clone().long().view(m.shape[-2], m.shape[-1]) for m in ref_m],
                               prev_m=None if prev_m is None else prev_m.clone().long().view(prev_m.shape[-2], prev_m.shape[-1])))
        t = self.t
        self.t += 1
        emb = torch.full((1, 4, 2, 2), float(t))
        return (None if prev_e is None else self.probs[t:t + 1].clone()), emb, memory


def test_run_sequence_matches_the_reference_eval_loop():
    """tests/golden/eval_loop_trace.pt holds what the REFERENCE's own `Evaluator.evaluating()`
    (networks/engine/eval_manager_mm.py:160-394, run unmodified by tools/make_eval_loop_golden.py) handed to the model
    on every call and the label maps it saved, for four synthetic sequences (absent id, id joining with ground truth,
    join on a memory frame, no candidate pool).  run_sequence must reproduce all of it bit for bit."""
    import os
    cases = torch.load(os.path.join(os.path.dirname(__file__), "golden", "eval_loop_trace.pt"))
    assert len(cases) == 4
    for c in cases:
        stub = _Replay(c["probs"])
        later = {t: lab for t, lab in c["labels"].items() if t != 0}
        preds = run_sequence(stub, torch.zeros(c["T"], 3, c["H"], c["W"]), c["labels"][0], c["K"],
                             mem_every=c["mem_every"], unc_ratio=c["unc_ratio"], later_labels=later)
        assert len(preds) == len(c["saved"]) == c["T"] - 1
        for t, (a, b) in enumerate(zip(preds, c["saved"])):
            assert torch.equal(a.long(), b), ("saved label map", c["seed"], t + 1)
        assert len(stub.calls) == len(c["calls"])
        for t, (a, b) in enumerate(zip(stub.calls, c["calls"])):
            assert a["n_ref"] == b["n_ref"], ("bank length", c["seed"], t)
            assert (a["prev_m"] is None) == (b["prev_m"] is None)
            if a["prev_m"] is not None:
                assert torch.equal(a["prev_m"], b["prev_m"]), ("previous mask", c["seed"], t)
            assert len(a["ref_m"]) == len(b["ref_m"])
            for i, (x, y) in enumerate(zip(a["ref_m"], b["ref_m"])):
                assert torch.equal(x, y), ("bank label map", c["seed"], t, i)


class _ReplayTTA:
    """the per-augmentation recording stub of tools/make_eval_loop_golden.py (call n = frame n // A, augmentation n % A)"""

    def __init__(self, probs, A):
        self.probs, self.A, self.calls, self.t = probs, A, [], 0

    def forward_for_eval(self, memory, ref_e, ref_m, prev_e, prev_m, img, pred_size=None, gt_ids=None):
        self.calls.append(dict(n_ref=len(ref_e), ref_m=[m.clone().long().view(m.shape[-2], m.shape[-1]) for m in ref_m],
                               prev_m=None if prev_m is None else prev_m.clone().long().view(prev_m.shape[-2], prev_m.shape[-1]),
                               img_hw=tuple(img.shape[-2:])))
        n = self.t
        self.t += 1
        emb = torch.full((1, 4, 2, 2), float(n))
        return (None if prev_e is None else self.probs[n // self.A, n % self.A][None].clone()), emb, memory


def test_run_sequence_tta_matches_the_reference_eval_loop():
    """tests/golden/eval_loop_tta_trace.pt: the REFERENCE's own loop (eval_manager_mm.py:160-394, unmodified) with
    test-time augmentation -- flip only; flip with an object joining at a ground-truth frame; two scales x flip with a
    join on a memory frame; three scales without flip.  run_sequence_tta must hand the model the same label maps on every
    call of every augmentation stream (incl. the unflipped confident mask in the mirrored stream's bank, the last
    augmentation's entropy, the seen-label list growing between augmentations) and save the same label maps."""
    import os
    import torch.nn.functional as F
    from aocb200.sequence import run_sequence_tta
    cases = torch.load(os.path.join(os.path.dirname(__file__), "golden", "eval_loop_tta_trace.pt"))
    assert len(cases) == 4
    for c in cases:
        augs = [(sz, fl) for sz in c["sizes"] for fl in ((False, True) if c["flip"] else (False,))]
        A = len(augs)
        samples = []
        for t in range(c["T"]):
            row = []
            for (h, w), fl in augs:
                lab = c["labels"].get(t)
                if lab is not None:
                    if (h, w) != (c["H"], c["W"]):
                        lab = F.interpolate(lab[None, None].float(), size=(h, w), mode="nearest")[0, 0].long()
                    if fl:
                        lab = torch.flip(lab, dims=[1])
                row.append(dict(img=torch.zeros(1, 3, h, w), label=lab, flip=fl, size=(c["H"], c["W"])))
            samples.append(row)
        stub = _ReplayTTA(c["probs"], A)
        preds = run_sequence_tta(stub, samples, c["K"], mem_every=c["mem_every"], unc_ratio=c["unc_ratio"])
        assert len(preds) == len(c["saved"]) == c["T"] - 1
        for t, (a, b) in enumerate(zip(preds, c["saved"])):
            assert torch.equal(a.long(), b), ("saved label map", c["seed"], t + 1)
        assert len(stub.calls) == len(c["calls"]) == c["T"] * A
        for n, (a, b) in enumerate(zip(stub.calls, c["calls"])):
            assert a["n_ref"] == b["n_ref"], ("bank length", c["seed"], n)
            assert a["img_hw"] == b["img_hw"]
            assert (a["prev_m"] is None) == (b["prev_m"] is None)
            if a["prev_m"] is not None:
                assert torch.equal(a["prev_m"], b["prev_m"]), ("previous mask", c["seed"], n)
            assert len(a["ref_m"]) == len(b["ref_m"]), ("bank masks", c["seed"], n)
            for i, (x, y) in enumerate(zip(a["ref_m"], b["ref_m"])):
                assert torch.equal(x, y), ("bank label map", c["seed"], n, i)
