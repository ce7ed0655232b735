import glob, os, sys
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aocb200.model import get_module
from aocb200.params import synthetic_state_dict
from aocb200.sequence import run_sequence
from aocb200.synth import make_clip

sd = synthetic_state_dict(1234)
m = get_module()(None, None); m.load_state_dict(sd); m = m.cuda().eval()
class Hooked:
    def __init__(s, m): s.m, s.logits, s.probs = m, [], []
    def forward_for_eval(s, *a, **k):
        out = s.m.forward_for_eval(*a, **k)
        if out[0] is not None:
            s.logits.append(s.m.engine().last_logits.clone().cpu()); s.probs.append(out[0].clone().cpu())
        return out
for path in sorted(glob.glob("tests/golden/tiny_*.pt")):
    if path.endswith("_fp64.pt"): continue
    g = torch.load(path); t64 = torch.load(path[:-3] + "_fp64.pt")
    frames, labels = make_clip(g["seed"], g["H"], g["W"], g["K"], g["T"])
    first = labels[0].clone()
    if g["drop"] is not None: first[first == g["drop"]] = 0
    hk = Hooked(m); np.random.seed(g["seed"])
    preds = run_sequence(hk, frames, first, g["K"], mem_every=g["mem_every"], unc_ratio=1.0, device=torch.device("cuda:0"))
    print(path, "keys", list(g.keys()))
    for t in range(len(preds)):
        mism = preds[t].cpu().to(torch.uint8) != g["preds"][t]
        if not mism.any(): continue
        ys, xs = torch.nonzero(mism, as_tuple=True)
        truth = t64["logits_fp64"][t]
        up = F.interpolate(truth, size=(g["H"], g["W"]), mode="bilinear", align_corners=True)[0]
        upe = F.interpolate(hk.logits[t].double(), size=(g["H"], g["W"]), mode="bilinear", align_corners=True)[0]
        upr = F.interpolate(g["logits"][t].double(), size=(g["H"], g["W"]), mode="bilinear", align_corners=True)[0]
        print(" frame", t + 1, "mismatches", int(mism.sum()))
        for y, x in list(zip(ys.tolist(), xs.tolist()))[:12]:
            print("   (%d,%d) ours=%d ref=%d truth_up=%s engine_up=%s ref_up=%s probs=%s" % (
                y, x, int(preds[t][y, x]), int(g["preds"][t][y, x]),
                np.round(up[:, y, x].numpy(), 3), np.round(upe[:, y, x].numpy(), 3), np.round(upr[:, y, x].numpy(), 3),
                np.round(hk.probs[t][0, :, y, x].numpy(), 4)))
