"""Compares the dumps of tools/mode_logits.py pairwise (run on the GPU box: the dumps are large)."""
import sys

import torch

L = {p.split("dbg_")[1][:-3]: torch.load(p) for p in sys.argv[1:]}
ks = sorted(L)
for i, a in enumerate(ks):
    for b in ks[i + 1:]:
        A, B = L[a], L[b]
        n = min(A["labels"].numel(), B["labels"].numel())
        print("%s | %s: logits %.3e, bank rows %.3e (max |S| %.2f), k-means labels differ at %d of %d rows, centroids %.3e, "
              "global %.3e cluster %.3e local %.3e" % (
                  a, b, (A["logits"] - B["logits"]).abs().max().item(), (A["S"] - B["S"]).abs().max().item(),
                  A["S"].abs().max().item(), (A["labels"][:n] != B["labels"][:n]).sum().item(), n,
                  (A["cent"] - B["cent"]).abs().max().item(), (A["g"] - B["g"]).abs().max().item(),
                  (A["gc"] - B["gc"]).abs().max().item(), (A["loc"] - B["loc"]).abs().max().item()))
