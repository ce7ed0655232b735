"""Per-sequence driver: the reference eval loop's bookkeeping around `forward_for_eval`.

A compact restatement of networks/engine/eval_manager_mm.py:172-361 for the single-scale, no-flip
case: first frame carries the ground-truth label, every later frame is predicted from (memory bank,
previous frame).  Works with anything exposing the reference's `forward_for_eval` contract
(aocnet.py:84): the reference model, the CPU oracle, and the CUDA engine.
"""
import math

import torch


def shannon_entropy(probs):
    # networks/layers/shannon_entropy.py:10-13
    return -1.0 * torch.sum(probs * torch.log(probs + 1e-6), dim=1, keepdim=True)


def run_sequence(model, frames, first_label, num_objects, mem_every=5, unc_ratio=1.0,
                 device=None, on_frame=None, keep_probs=False):
    """frames [T,3,H,W] (normalised), first_label [H,W] ints 0..K.  Returns list of predicted label
    maps [H,W] (int64) for frames 1..T-1 (and the per-frame probabilities when keep_probs)."""
    T, _, H, W = frames.shape
    dev = device if device is not None else frames.device
    gt_ids = torch.tensor([num_objects], device=dev)
    ref_emb, ref_mask_conf = [], []          # eval_manager_mm.py:184-189
    prev_emb = prev_mask = None
    memory = [[None, None]]                   # :168-170,:205-207 (BLOCK_NUM = 2 placeholders)
    preds, probs_out = [], []
    seen = set(int(v) for v in torch.unique(first_label).tolist())            # label_all_list :262-265
    for t in range(T):
        img = frames[t:t + 1].to(dev, non_blocking=True)
        probs, emb, memory = model.forward_for_eval(
            memory, ref_emb, ref_mask_conf, prev_emb, prev_mask, img,
            pred_size=[H, W], gt_ids=gt_ids)                                   # :246-249
        if t == 0:
            lab = first_label.to(dev).view(1, 1, H, W)
            ref_emb.append(emb); ref_mask_conf.append(lab)                     # :275-281
            prev_emb, prev_mask = emb, lab
            continue
        exist = [i for i in range(probs.shape[1]) if i in seen]                # :252-261
        if len(exist) != probs.shape[1]:
            keep = torch.zeros(probs.shape[1], device=probs.device)
            keep[exist] = 1.0
            probs_exist = probs[:, exist]
            probs = probs * keep.view(1, -1, 1, 1)
        else:
            probs_exist = probs
        pred = torch.argmax(probs[0], dim=0)                                   # :318-320
        cur = pred.view(1, 1, H, W)
        if mem_every > -1 and t % mem_every == 0:                              # :309-312,:356-361
            unc = shannon_entropy(probs_exist)[0, 0]
            region = (unc > unc_ratio).long()
            conf = (pred * (1 - region) + 125 * region).view(1, 1, H, W)
            ref_emb.append(emb); ref_mask_conf.append(conf)
        prev_emb, prev_mask = emb, cur                                         # :315,:351-354
        preds.append(pred)
        if keep_probs:
            probs_out.append(probs)
        if on_frame is not None:
            on_frame(t, probs, pred)
    return (preds, probs_out) if keep_probs else preds
