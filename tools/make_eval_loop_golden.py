"""Pins aocb200/sequence.py::run_sequence (the restatement of the reference's eval-loop bookkeeping that the device-resident
path is checked against) to the REFERENCE's own loop: runs `Evaluator.evaluating()` of
networks/engine/eval_manager_mm.py -- unmodified, imported from /root/reference -- on a synthetic sequence with a
recording stub in place of the network, and stores what the loop handed to the model on every call (memory-bank label
maps incl. the label-125 "confident" masks, previous masks, bank length) and the label maps it saved.
tests/test_sequence_cpu.py replays run_sequence on the same stub outputs and compares.  Build container only.

    python tools/make_eval_loop_golden.py          # -> tests/golden/eval_loop_trace.pt, eval_loop_tta_trace.pt

The second file pins aocb200/sequence.py::run_sequence_tta (TEST_FLIP / TEST_MULTISCALE, eval_manager_mm.py:195-361) the
same way: the dataset hands the loop several augmented samples per frame (plain / mirrored, two scales), the stub
returns different probabilities per augmentation, and every label map the loop handed back or saved is recorded.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "eval_loop_trace.pt")


def stub_probs(seed, T, O, H, W):
    """softmax maps with entropies on both sides of the threshold; frame t uses probs[t] (probs[0] unused)"""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(T, O, H, W, generator=g) * 2.0
    return torch.softmax(logits, dim=1)


class RecordingStub:
    def __init__(self, probs):
        self.probs, self.calls, self.t = probs, [], 0

    def eval(self):
        return self

    def forward_for_eval(self, memory, ref_e, ref_m, prev_e, prev_m, img, pred_size=None, gt_ids=None):
        self.calls.append(dict(n_ref=len(ref_e), ref_m=[m.clone().long().view(m.shape[-2], m.shape[-1]) for m in ref_m],
                               prev_m=None if prev_m is None else prev_m.clone().long().view(prev_m.shape[-2], prev_m.shape[-1])))
        t = self.t
        self.t += 1
        emb = torch.full((1, 4, 2, 2), float(t))
        return (None if prev_e is None else self.probs[t:t + 1].clone()), emb, memory


class SeqDataset(torch.utils.data.Dataset):
    """what DAVIS_Test / YOUTUBE_VOS_Test hand to the loop after MultiRestrictSize + MultiToTensor (datasets_m.py:458-494)"""

    def __init__(self, name, T, H, W, labels):
        self.seq_name, self.T, self.H, self.W, self.labels = name, T, H, W, labels

    def __len__(self):
        return self.T

    def __getitem__(self, i):
        s = {"current_img": torch.zeros(3, self.H, self.W)}
        if i in self.labels:
            s["current_label"] = self.labels[i].to(torch.uint8).view(1, self.H, self.W)
        s["meta"] = {"seq_name": self.seq_name, "frame_num": self.T, "obj_num": 3, "obj_list": [0, 1, 2, 3],
                     "current_name": "%05d.jpg" % i, "height": self.H, "width": self.W, "flip": False}
        return [s]


class TTAStub(RecordingStub):
    """per-augmentation probabilities: call n = frame n // A, augmentation n % A"""

    def __init__(self, probs, A):
        super().__init__(probs)
        self.A = A

    def forward_for_eval(self, memory, ref_e, ref_m, prev_e, prev_m, img, pred_size=None, gt_ids=None):
        self.calls.append(dict(n_ref=len(ref_e), ref_m=[m.clone().long().view(m.shape[-2], m.shape[-1]) for m in ref_m],
                               prev_m=None if prev_m is None else prev_m.clone().long().view(prev_m.shape[-2], prev_m.shape[-1]),
                               img_hw=tuple(img.shape[-2:])))
        n = self.t
        self.t += 1
        emb = torch.full((1, 4, 2, 2), float(n))
        return (None if prev_e is None else self.probs[n // self.A, n % self.A][None].clone()), emb, memory


def resize_label(lab, size):
    return torch.nn.functional.interpolate(lab[None, None].float(), size=size, mode="nearest")[0, 0].long()


def tta_augs(c):
    """(scale size, flip) per augmentation in the order of MultiRestrictSize(+flip): custom_transforms.py:433-462"""
    return [(sz, fl) for sz in c["sizes"] for fl in ((False, True) if c["flip"] else (False,))]


class TTADataset(SeqDataset):
    def __init__(self, name, c):
        super().__init__(name, c["T"], c["H"], c["W"], c["labels"])
        self.c = c

    def __getitem__(self, i):
        out = []
        for (h, w), fl in tta_augs(self.c):
            s = {"current_img": torch.zeros(3, h, w)}
            if i in self.labels:
                lab = self.labels[i]
                if (h, w) != (self.H, self.W):
                    lab = resize_label(lab, (h, w))
                if fl:
                    lab = torch.flip(lab, dims=[1])
                s["current_label"] = lab.to(torch.uint8).view(1, h, w)
            s["meta"] = {"seq_name": self.seq_name, "frame_num": self.T, "obj_num": 3, "obj_list": [0, 1, 2, 3],
                         "current_name": "%05d.jpg" % i, "height": self.H, "width": self.W, "flip": fl}
            out.append(s)
        return out


def tta_case(seed, T, H, W, K, mem_every, unc_ratio, absent, join_at, sizes, flip):
    c = case(seed, T, H, W, K, mem_every, unc_ratio, absent, join_at)
    A = len(sizes) * (2 if flip else 1)
    g = torch.Generator().manual_seed(seed + 7)
    c["probs"] = torch.softmax(torch.randn(T, A, K + 1, H, W, generator=g) * 2.0, dim=2)
    c["sizes"], c["flip"] = sizes, flip
    return c


def case(seed, T, H, W, K, mem_every, unc_ratio, absent, join_at):
    g = torch.Generator().manual_seed(seed + 100)
    first = torch.randint(0, K + 1, (H, W), generator=g)
    first[first == absent] = 0
    labels = {0: first}
    if join_at is not None:
        j = torch.zeros(H, W, dtype=torch.long)
        j[torch.rand(H, W, generator=g) < 0.2] = absent
        labels[join_at] = j
    return dict(seed=seed, T=T, H=H, W=W, K=K, mem_every=mem_every, unc_ratio=unc_ratio, labels=labels,
                probs=stub_probs(seed, T, K + 1, H, W))


def run_reference_loop(c, tta=False):
    load_reference()
    sys.modules["matplotlib.pyplot"].rcParams = {}
    em = importlib.import_module("networks.engine.eval_manager_mm")
    saved = []
    em.save_mask = lambda lab, path: saved.append(lab.clone().long())
    em.zip_folder = lambda *a, **k: None
    ev = object.__new__(em.Evaluator)
    ev.cfg = types.SimpleNamespace(BLOCK_NUM=2, TEST_WORKERS=0)
    if tta:
        ev.model = TTAStub(c["probs"], len(tta_augs(c)))
        ev.dataset = [TTADataset("seq", c)]
    else:
        ev.model = RecordingStub(c["probs"])
        ev.dataset = [SeqDataset("seq", c["T"], c["H"], c["W"], c["labels"])]
    ev.mem_every, ev.unc_ratio, ev.gpu = c["mem_every"], c["unc_ratio"], 0
    ev.result_root = ev.source_folder = "/tmp/eval_loop_golden"
    ev.zip_dir = "/tmp/eval_loop_golden.zip"
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self           # the loop moves every sample to its GPU; this box has none
    try:
        ev.evaluating()
    finally:
        torch.Tensor.cuda = cuda
    return ev.model.calls, saved


def main():
    cases = [case(1, 7, 6, 8, 3, 2, 0.8, 2, None),           # one id never appears: label-existence filter on every frame
             case(2, 8, 6, 8, 3, 2, 0.8, 2, 3),              # the id joins with ground truth at frame 3
             case(3, 7, 5, 7, 3, 3, 0.9, 3, 3),              # join on a memory frame (t % mem_every == 0)
             case(4, 6, 5, 7, 2, -1, 1.0, 9, None)]          # mem_every = -1: no confident candidate pool
    out = []
    for c in cases:
        calls, saved = run_reference_loop(c)
        assert len(calls) == c["T"] and len(saved) == c["T"] - 1
        n125 = sum(int((m == 125).sum()) for call in calls for m in call["ref_m"])
        print("case seed %d: %d calls, bank length at the end %d, label-125 pixels handed to the model %d"
              % (c["seed"], len(calls), calls[-1]["n_ref"], n125))
        out.append(dict(c, calls=calls, saved=saved))
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    tta = [tta_case(11, 7, 6, 8, 3, 2, 0.8, 2, None, [(6, 8)], True),                 # flip only, an id never appears
           tta_case(12, 8, 6, 8, 3, 2, 0.8, 2, 3, [(6, 8)], True),                    # flip, the id joins with ground truth at frame 3
           tta_case(13, 7, 6, 8, 3, 3, 0.9, 3, 3, [(6, 8), (9, 11)], True),           # two scales x flip, join on a memory frame
           tta_case(14, 6, 5, 7, 2, 2, 0.7, 9, None, [(5, 7), (7, 9), (4, 6)], False)]  # three scales, no flip
    out = []
    for c in tta:
        calls, saved = run_reference_loop(c, tta=True)
        A = len(tta_augs(c))
        assert len(calls) == c["T"] * A and len(saved) == c["T"] - 1
        n125 = sum(int((m == 125).sum()) for call in calls for m in call["ref_m"])
        print("tta case seed %d: %d augmentations, %d calls, label-125 pixels handed to the model %d"
              % (c["seed"], A, len(calls), n125))
        out.append(dict(c, calls=calls, saved=saved))
    out_tta = os.path.join(ROOT, "tests", "golden", "eval_loop_tta_trace.pt")
    torch.save(out, out_tta)
    print("wrote", out_tta, os.path.getsize(out_tta), "bytes")


if __name__ == "__main__":
    main()
