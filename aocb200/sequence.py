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
                 device=None, on_frame=None, keep_probs=False, later_labels=None):
    """frames [T,3,H,W] (normalised), first_label [H,W] ints 0..K.  Returns list of predicted label
    maps [H,W] (int64) for frames 1..T-1 (and the per-frame probabilities when keep_probs).
    later_labels: {t: [H,W] label map} ground truth given at later frames (YouTube-VOS: objects that
    enter after the first frame, eval_manager_mm.py:288-292,:321-349)."""
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
        join = later_labels.get(t) if later_labels else None
        if join is not None:                                                   # :321-349 new objects join here
            join = join.to(pred.device).long()
            seen |= set(int(v) for v in torch.unique(join).tolist())           # :262-265
            keep = (join == 0).long()
            pred = pred * keep + join * (1 - keep)
            unc = shannon_entropy(probs_exist)[0, 0] * keep                    # (join < 0) is empty for label maps
            region = (unc > unc_ratio).long()
            conf = (pred * (1 - region) + 125 * region).view(1, 1, H, W)
            ref_emb.append(emb); ref_mask_conf.append(conf)                    # :296-297,:349
        cur = pred.view(1, 1, H, W)
        if join is None and mem_every > -1 and t % mem_every == 0:             # :309-312,:356-361
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


def flip_w(t):
    """utils/image.py:52-55 on the last (width) axis"""
    return torch.flip(t, dims=[t.dim() - 1])


def run_sequence_tta(model, samples, num_objects, mem_every=5, unc_ratio=1.0, device=None, on_frame=None):
    """The reference eval loop WITH test-time augmentation (TEST_FLIP / TEST_MULTISCALE; eval_manager_mm.py:195-361):
    every frame arrives as a list of augmented samples -- what MultiRestrictSize + MultiToTensor produce
    (custom_transforms.py:387-487): per scale the plain and, with flipping, the mirrored image -- each augmentation keeps
    its own memory bank / previous frame / decoder memory, the probabilities (all at the ORIGINAL size, mirrored ones
    flipped back, :304) are averaged, and the averaged label map feeds every stream's bank and previous mask.

    samples[t] = list over augmentations of dict(img [1,3,h_a,w_a], label [h_a,w_a] ints or None, flip bool); the
    original size is samples[t][i]['size'] = (H, W) (default: the image size of augmentation 0).
    Returns the list of label maps [H,W] for frames 1..T-1.

    The loop is restated with its three order-dependent behaviours, because a drop-in has to hand the model exactly what
    the reference hands it (pinned by tests/golden/eval_loop_tta_trace.pt, recorded from the reference's own loop):
      1. the seen-label list is updated BETWEEN the augmentations of a ground-truth frame (:262-265 sits inside the
         augmentation loop, after the filter of that augmentation);
      2. the entropy map that decides label 125 is the LAST augmentation's -- computed from its probabilities before they
         are flipped back, so in mirrored coordinates when that augmentation is mirrored (:266-270, :304, :339);
      3. on a memory frame EVERY stream's confident bank mask is the unflipped map (:356-361), while a mirrored stream's
         previous mask (and, on a ground-truth frame, its bank mask) is the flipped one (:343-345, :351-354).
    A ground-truth label on a non-mirrored augmentation of a different size than the original cannot be joined by the
    reference either (:322-325 mixes the sizes); that case raises here."""
    dev = device
    A = len(samples[0])
    gt_ids = torch.tensor([num_objects], device=dev) if dev is not None else torch.tensor([num_objects])
    ref_emb = [[] for _ in range(A)]
    ref_conf = [[] for _ in range(A)]
    prev_emb, prev_mask = [None] * A, [None] * A
    memory = [[[None, None]] for _ in range(A)]
    seen = []                                                                  # label_all_list
    preds = []
    for t, augs in enumerate(samples):
        assert len(augs) == A
        H, W = augs[0].get("size", tuple(augs[0]["img"].shape[-2:]))
        all_preds, join, update = [], None, False
        exist_probs = unc = None
        for a, smp in enumerate(augs):
            img = smp["img"] if dev is None else smp["img"].to(dev, non_blocking=True)
            lab = smp.get("label")
            if lab is not None:
                lab = lab if dev is None else lab.to(dev)
                lab4 = lab.view(1, 1, lab.shape[-2], lab.shape[-1])
            probs, emb, memory[a] = model.forward_for_eval(memory[a], ref_emb[a], ref_conf[a], prev_emb[a], prev_mask[a],
                                                           img, pred_size=[H, W], gt_ids=gt_ids)            # :246-249
            if probs is not None:                                                                        # :252-261
                keep = torch.zeros(probs.shape[1], device=probs.device)
                ex = [i for i in range(probs.shape[1]) if i in seen]
                keep[ex] = 1.0
                exist_probs = probs[:, ex]
                probs = probs * keep.view(1, -1, 1, 1)
            if lab is not None:                                                                          # :262-265
                for v in torch.unique(lab).tolist():
                    if int(v) not in seen:
                        seen.append(int(v))
            if t == 0:                                                                                   # :267-277
                assert lab is not None, "the first frame carries the ground-truth label"
                ref_emb[a].append(emb); ref_conf[a].append(lab4)
                prev_emb[a], prev_mask[a] = emb, lab4
                continue
            if smp["flip"]:
                probs = flip_w(probs)                                                                    # :279-280
            if not smp["flip"] and lab is not None and join is None:                                     # :283-284
                join = lab
            all_preds.append(probs)
            if lab is not None:
                ref_emb[a].append(emb)                                                                   # :290-291
            else:
                unc = shannon_entropy(exist_probs)                                                       # :300 (this augmentation's)
                if mem_every > -1 and t % mem_every == 0:                                                # :303-306
                    ref_emb[a].append(emb)
                    update = True
            prev_emb[a] = emb
        if t == 0:
            continue
        mean = torch.mean(torch.cat(all_preds, dim=0), dim=0)                                            # :312-314
        pred = torch.argmax(mean, dim=0)
        if join is not None:                                                                             # :315-319
            if tuple(join.shape[-2:]) != (H, W):
                raise ValueError("a ground-truth label at an augmented size cannot be joined (eval_manager_mm.py:322-325)")
            join = join.to(pred.device).long().view(H, W)
            keepj = (join == 0).long()
            pred = pred * keepj + join * (1 - keepj)
        cur = pred.view(1, 1, H, W)
        flipped = flip_w(pred).view(1, 1, H, W) if augs[-1]["flip"] else None                            # :321-323 (LAST augmentation)
        for a, smp in enumerate(augs):
            if join is not None:                                                                         # :326-343
                if smp["flip"]:
                    ref_conf[a].append(flipped)
                else:
                    u = shannon_entropy(exist_probs)[0, 0] * keepj
                    unc = u.view(1, 1, H, W)
                    region = (u > unc_ratio).long()
                    ref_conf[a].append((pred * (1 - region) + 125 * region).view(1, 1, H, W))
            prev_mask[a] = flipped if smp["flip"] else cur                                               # :345-348
            if update:                                                                                   # :350-355
                region = (unc.view(H, W) > unc_ratio).long()
                ref_conf[a].append((pred * (1 - region) + 125 * region).view(1, 1, H, W))
        preds.append(pred)
        if on_frame is not None:
            on_frame(t, mean, pred)
    return preds


class DeviceSequence:
    """The same bookkeeping with every piece of per-sequence state resident on the GPU: label maps
    are uint8 device tensors, the label-existence filter (eval_manager_mm.py:252-270), the argmax (:318-320) and the
    entropy -> label-125 "confident" mask (:339-349, shannon_entropy.py:10-13) come out of the engine's fused
    upsample + softmax kernel (`aoc_upsample_softmax_label_f32`), so a predicted frame adds no torch kernels and
    no host read-back to `forward_for_eval`.  CUDA engine only (`model.engine()`); no fallback.
    The sequence's settings (seen-label word, entropy threshold, probability output off unless `keep_probs`) are applied
    to the engine for the duration of each step() and undone afterwards: plain `forward_for_eval` callers and other
    sequences sharing the model never see them."""

    def __init__(self, model, num_objects, mem_every=5, unc_ratio=1.0, device=None, keep_probs=False):
        self.m, self.eng = model, model.engine()
        self.dev = device if device is not None else self.eng.dev
        self.K, self.mem_every = int(num_objects), int(mem_every)
        self.unc_ratio, self.keep_probs = float(unc_ratio), bool(keep_probs)
        self.gt_ids = self.K                                                    # a Python int: no device read per frame
        self.ref_e, self.ref_m = [], []
        self.prev_e = self.prev_m = None
        self.memory = [[None, None]]
        self.seen = set()
        self.bits = 0                                                           # set by the first ground-truth frame
        self.probs = None
        self.t = -1

    def _see(self, label):
        new = set(int(v) for v in torch.unique(label).tolist()) - self.seen      # GT frames only: one small read-back
        if new:
            self.seen |= new
            self.bits = self.eng.seen_bits(self.seen)

    def _forward(self, img):
        eng = self.eng
        saved = (eng.unc_ratio, eng.want_probs)
        eng.unc_ratio, eng.want_probs = self.unc_ratio, self.keep_probs
        eng._set_exist_bits(self.bits)
        try:
            return self.m.forward_for_eval(self.memory, self.ref_e, self.ref_m, self.prev_e, self.prev_m, img,
                                           pred_size=[int(img.shape[-2]), int(img.shape[-1])], gt_ids=self.gt_ids)
        finally:
            eng.unc_ratio, eng.want_probs = saved
            eng._set_exist_bits(-1)

    def step(self, img, gt_label=None):
        """img [1,3,H,W] normalised frame (device or pinned host); gt_label [H,W] ints if this frame carries ground
        truth (always the first).  Returns the frame's label map, uint8 [H,W] on the device (the GT itself for t = 0)."""
        self.t += 1
        t = self.t
        H, W = int(img.shape[-2]), int(img.shape[-1])
        img = img.to(self.dev, non_blocking=True)
        if gt_label is not None:
            gt_label = gt_label.to(self.dev).to(torch.uint8).view(H, W)
            if t == 0:
                self._see(gt_label)
        probs, emb, self.memory = self._forward(img)
        if t == 0:
            assert gt_label is not None, "the first frame carries the ground-truth label"
            lab = gt_label.view(1, 1, H, W)
            self.ref_e.append(emb); self.ref_m.append(lab)
            self.prev_e, self.prev_m = emb, lab
            return gt_label
        self.probs = probs
        pred = self.eng.last_label.clone()
        if gt_label is not None:                                               # new objects join (:321-349)
            self._see(gt_label)
            conf = torch.where(gt_label != 0, gt_label, self.eng.last_conf_label)
            pred = torch.where(gt_label != 0, gt_label, pred)
            self.ref_e.append(emb); self.ref_m.append(conf.view(1, 1, H, W))
        elif self.mem_every > -1 and t % self.mem_every == 0:
            self.ref_e.append(emb); self.ref_m.append(self.eng.last_conf_label.clone().view(1, 1, H, W))
        self.prev_e, self.prev_m = emb, pred.view(1, 1, H, W)
        return pred


def run_sequence_device(model, frames, first_label, num_objects, mem_every=5, unc_ratio=1.0, later_labels=None):
    """run_sequence() on a DeviceSequence: same arguments, label maps returned as uint8 device tensors."""
    seq = DeviceSequence(model, num_objects, mem_every, unc_ratio)
    seq.step(frames[0:1], first_label)
    return [seq.step(frames[t:t + 1], later_labels.get(t) if later_labels else None)
            for t in range(1, frames.shape[0])]
