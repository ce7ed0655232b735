"""Parity at BASELINE.json's full sizes, through the reference-facing API (get_module() -> forward_for_eval).

 * configs[1]/[2]  480p (481x849), 5 objects: against the CPU oracle on the same clip / weights / RNG stream (the oracle
   needs a few seconds per frame at this size, so two predicted frames with the bank growing every frame).
 * configs[3]      720p, 10 objects and configs[4] 1080p, 5 objects with a growing bank: size-independent properties --
   probabilities are a softmax (finite, in [0, 1], sum to 1), objects absent from every ground truth get no pixel,
   the run is bit-reproducible from the numpy seed, the CUDA-graph schedule and the plain-launch schedule are the same
   arithmetic bit for bit, and the split-K convolution schedule changes logits by fp32 rounding only.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(state_dict):
    from aocb200.model import get_module
    m = get_module()(None, None)
    m.load_state_dict(state_dict)
    return m.cuda().eval()


def _run(model, frames, first, K, seed, mem_every, graphs=True, opts=()):
    from aocb200.sequence import run_sequence
    eng = model.engine()
    old = eng.use_graphs
    eng.use_graphs = graphs
    for k, v in opts:
        eng.L.set_option(k, v)
    logits = []

    def on_frame(t, probs, pred):
        logits.append(eng.last_logits.clone())
    np.random.seed(seed)
    try:
        preds, probs = run_sequence(model, frames, first, K, mem_every=mem_every, device=torch.device("cuda:0"),
                                    on_frame=on_frame, keep_probs=True)
    finally:
        eng.use_graphs = old
        for k, v in opts:
            eng.L.set_option(k, 1)
    torch.cuda.synchronize()
    return preds, probs, logits


def test_480p_k5_vs_oracle(model, state_dict):
    """Engine and CPU oracle frame by frame on a 481x849, 5-object clip.  Both are fed the ORACLE's label maps (bank
    masks, previous-frame mask): the k-means initial rows are drawn with np.random.choice(N_i, k) from the per-object
    pixel counts N_i, so a single argmax flip at a numerical tie in frame 1 would re-draw every proxy of frame 2 --
    a property of the reference's RNG coupling, not of either implementation.  With equal masks the draws are equal
    and the logits must agree to fp32 rounding."""
    import os
    from aocb200.synth import make_clip, restrict_size
    from oracle.aoc_oracle import AOCOracle
    K = 5
    H, W = restrict_size(480, 854)
    assert (H, W) == (481, 849)
    frames, labels = make_clip(3, H, W, K, 3)
    dev = torch.device("cuda:0")
    orc = AOCOracle(state_dict)
    eng = model.engine()
    gt_c, gt_g = torch.tensor([K]), torch.tensor([K], device=dev)
    truth = torch.load(os.path.join(os.path.dirname(__file__), "golden", "full480_k5_fp64.pt"))
    with torch.no_grad():
        _, eo, mo = orc.forward_for_eval([[None, None]], [], [], None, None, frames[0:1], [H, W], gt_c)
        _, ee, me = model.forward_for_eval([[None, None]], [], [], None, None, frames[0:1].to(dev), [H, W], gt_g)
        lab = labels[0].view(1, 1, H, W).long()
        refs_o, refs_e, masks_c, masks_g = [eo], [ee], [lab], [lab.to(dev)]
        prev_o, prev_e, prev_c, prev_g = eo, ee, lab, lab.to(dev)
        for t in (1, 2):
            np.random.seed(3 if t == 1 else 100 + t)
            po, eo, mo = orc.forward_for_eval(mo, refs_o, masks_c, prev_o, prev_c, frames[t:t + 1], [H, W], gt_c)
            lo = orc.last_logits.clone()
            np.random.seed(3 if t == 1 else 100 + t)
            pe, ee, me = model.forward_for_eval(me, refs_e, masks_g, prev_e, prev_g, frames[t:t + 1].to(dev), [H, W], gt_g)
            le = eng.last_logits.clone().cpu()
            ad = (le - lo).abs()
            d, q = ad.max().item(), torch.quantile(ad.flatten()[::7], 0.999).item()
            yo, ye = torch.argmax(po[0], 0), torch.argmax(pe[0], 0).cpu()
            eq = (yo == ye).float().mean().item()
            print("[parity] 480p K=5 frame %d (oracle masks fed to both): max|dlogit| %.3e, 99.9th percentile %.3e "
                  "(logit range %.1f), argmax-equal %.6f" % (t, d, q, lo.abs().max().item(), eq))
            if t == 1:
                # frame 1 also has a float64 evaluation (tools/make_fullsize_truth.py): same bar as the tiny fixtures
                t64, n32 = truth["logits_fp64"], truth["oracle32_noise"]
                d64 = (le.double() - t64).abs().max().item()
                print("[parity] 480p K=5 frame 1: |engine-fp64| %.3e vs |oracle fp32-fp64| %.3e" % (d64, n32))
                # observed x 1.5 (B200, round 2): |engine - fp64| 6.85e-3 (the oracle's own fp32 run: 1.69e-2), |engine - oracle|
                # 1.80e-2 (= the oracle's noise), argmax-equal 0.999628
                assert d64 <= 1.03e-2 and d64 <= n32, (d64, n32)
                assert d <= 2.7e-2, d
                assert eq >= 1.0 - 5.6e-4, eq
                mism = ye.to(torch.uint8) != truth["pred_fp32"]
                if mism.any():
                    import torch.nn.functional as F
                    up = F.interpolate(t64, size=(H, W), mode="bilinear", align_corners=True)[0]
                    top2 = torch.topk(up, 2, dim=0)[0]
                    assert (top2[0] - top2[1])[mism].max().item() <= 2.0 * (d64 + n32), "argmax differs away from a tie"
                    assert mism.float().mean().item() < 1e-3
            else:
                # no float64 evaluation for frame 2; observed x 1.5: 1.60e-2, argmax-equal 0.999650
                assert d <= 2.4e-2 and eq >= 1.0 - 5.3e-4, (d, q, eq)
            mask = yo.view(1, 1, H, W)
            refs_o.append(eo); refs_e.append(ee); masks_c.append(mask); masks_g.append(mask.to(dev))
            prev_o, prev_e, prev_c, prev_g = eo, ee, mask, masks_g[-1]


def test_480p_k5_vs_reference_fixture(model):
    """The engine against the REFERENCE ITSELF at BASELINE.json's headline configuration: tests/golden/full480_k5_ref.npz
    holds the logits and label maps the repaired reference (tools/ref_loader.py) produced on this clip / these weights /
    this numpy stream (tools/make_ref480_golden.py); frame 2 is teacher-forced with the reference's frame-1 label map (the
    k-means draws depend on per-object pixel counts), so it sees a two-frame bank and a filled decoder memory.  The
    oracle sits 3.6e-4 / 3.3e-4 from the same vectors (tests/test_oracle_golden.py)."""
    import os
    from aocb200.synth import make_clip
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "full480_k5_ref.npz"))
    seed, K, H, W, n_pred = (int(g[k]) for k in ("seed", "K", "H", "W", "n_pred"))
    frames, labels = make_clip(seed, H, W, K, n_pred + 1)
    dev = torch.device("cuda:0")
    eng = model.engine()
    gt = torch.tensor([K], device=dev)
    with torch.no_grad():
        _, emb, mem = model.forward_for_eval([[None, None]], [], [], None, None, frames[0:1].to(dev), [H, W], gt)
        lab = labels[0].view(1, 1, H, W).long().to(dev)
        refs, masks, prev_e, prev_m = [emb], [lab], emb, lab
        for t in range(1, n_pred + 1):
            np.random.seed(seed if t == 1 else 100 * seed + t)
            probs, emb, mem = model.forward_for_eval(mem, refs, masks, prev_e, prev_m, frames[t:t + 1].to(dev), [H, W], gt)
            ref_l = torch.from_numpy(g["logits"][t - 1])
            d = (eng.last_logits.reshape(K + 1, -1).cpu() - ref_l).abs().max().item()
            ref_pred = torch.from_numpy(g["preds"][t - 1])
            bad = int((torch.argmax(probs[0], 0).to(torch.uint8).cpu() != ref_pred).sum())
            print("[parity] engine vs REFERENCE, 480p K=5 frame %d: max|dlogit| %.3e (logit range %.1f), argmax differs at "
                  "%d of %d px" % (t, d, ref_l.abs().max().item(), bad, ref_pred.numel()))
            assert d <= REF480_BOUNDS[t - 1][0] and bad <= REF480_BOUNDS[t - 1][1], (t, d, bad)
            m = ref_pred.view(1, 1, H, W).long().to(dev)
            refs.append(emb); masks.append(m)
            prev_e, prev_m = emb, m


# (max |dlogit|, argmax pixels) per frame = observed x 1.5 (B200, round 2; the [parity] lines the test prints): observed
# 1.76e-2 / 127 px of 408 369 on frame 1 -- the reference's own fp32 result sits 1.69e-2 from a float64 evaluation of this
# frame and the engine 6.8e-3 (test_480p_k5_vs_oracle), so this IS the reference's rounding noise -- and 4.39e-2 / 146 px on
# frame 2, where the engine's own k-means runs on bank embeddings that differ from the CPU's in the sixth digit (boundary
# rows change cluster: a discrete step of the algorithm, see tools/make_cfg_truth.py)
REF480_BOUNDS = [(2.7e-2, 190), (6.6e-2, 220)]

# observed x 1.5 (B200, round 2): see the [parity] lines these tests print
CFG_BOUNDS = {
    # name: (|engine - fp64| with pinned proxies, |engine - oracle fp32| with pinned proxies, argmax pixels differing)
    # observed: cfg4 1.57e-2 / 1.63e-2 / 236 px of 923 601 (the oracle's own fp32 run: 9.8e-3 from float64)
    "cfg4_720p_k10": (2.4e-2, 2.5e-2, 360),
    # observed: cfg5 1.03e-2 / 1.11e-2 / 558 px of 2 061 233 (the oracle's own fp32 run: 7.4e-3 from float64, its argmax differs
    # from the float64 run's at 467 px)
    "cfg5_1080p_k5_bank3": (1.6e-2, 1.7e-2, 840),
}


@pytest.mark.parametrize("name", sorted(CFG_BOUNDS))
def test_cfg4_cfg5_vs_oracle_fixture(model, name):
    """BASELINE configs[3] (721x1281, 10 objects, frame 1) and configs[4] (1073x1921, 5 objects, frame 3 with a
    three-frame bank) against the CPU oracle's committed vectors (tools/make_cfg_truth.py: the oracle needs minutes per
    frame at these sizes).  The engine is fed the oracle's label maps and numpy seeds, and -- for the logit comparison --
    pinned to the oracle's k-means proxies of the compared frame: k-means is a discrete step whose boundary rows flip
    under 1e-6 perturbations (see the tool's docstring), which would otherwise drown a 1e-2 comparison in 0.1-0.3 jumps
    that are a property of the algorithm.  A second run with the engine's own k-means is held to argmax agreement."""
    import os
    import torch.nn.functional as F
    from aocb200.synth import make_clip
    path = os.path.join(os.path.dirname(__file__), "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated (python tools/make_cfg_truth.py)" % name)
    f = np.load(path)
    K, H, W, seed, n_pred = int(f["K"]), int(f["H"]), int(f["W"]), int(f["seed"]), int(f["n_pred"])
    frames, labels = make_clip(seed, H, W, K, n_pred + 1)
    dev = torch.device("cuda:0")
    eng = model.engine()
    gt = torch.tensor([K], device=dev)
    l32, l64 = torch.from_numpy(f["logits_fp32"]), torch.from_numpy(f["logits_fp64"]).double()
    n32 = float(f["oracle32_noise"])
    want = torch.from_numpy(f["pred_fp32"])
    res = {}
    old = eng.use_graphs
    eng.use_graphs = False                                  # the proxy pin is a plain-launch hook
    try:
        for pinned in (True, False):
            _, emb, mem = model.forward_for_eval([[None, None]], [], [], None, None, frames[0:1].to(dev), [H, W], gt)
            lab = labels[0].view(1, 1, H, W).to(dev)
            refs, masks, prev_e, prev_m = [emb], [lab], emb, lab
            for t in range(1, n_pred + 1):
                np.random.seed(seed if t == 1 else 100 * seed + t)
                # pinned: every predicted frame uses the oracle's proxies of that frame (earlier frames reach the
                # compared one through the decoder memory)
                eng.force_proxies = {k: f[k][t - 1] for k in ("prox_cen", "prox_avg", "prox_ncen", "prox_navg")} \
                    if pinned else None
                probs, emb, mem = model.forward_for_eval(mem, refs, masks, prev_e, prev_m, frames[t:t + 1].to(dev),
                                                         [H, W], gt)
                if t < n_pred:
                    m = torch.from_numpy(f["fed"][t - 1]).view(1, 1, H, W).to(dev)
                    refs.append(emb); masks.append(m)
                    prev_e, prev_m = emb, m
            le = eng.last_logits.clone().cpu()
            pred = torch.argmax(probs[0], 0).to(torch.uint8).cpu()
            res[pinned] = ((le.double() - l64).abs().max().item(), (le - l32).abs().max().item(), pred != want, le)
    finally:
        eng.force_proxies = None
        eng.use_graphs = old
    d64, d32, mism, le = res[True]
    print("[parity] %s (oracle proxies pinned): |engine-fp64| %.3e  |engine-oracle fp32| %.3e  |oracle fp32-fp64| %.3e  "
          "(logit range %.1f)  argmax differs from the oracle's at %d of %d px"
          % (name, d64, d32, n32, l64.abs().max().item(), int(mism.sum()), mism.numel()))
    b64, b32, bpx = CFG_BOUNDS[name]
    assert d64 <= b64 and d32 <= b32, (d64, d32)
    assert int(mism.sum()) <= bpx
    if mism.any():                                          # every differing pixel is a float64 near-tie
        up = F.interpolate(l64.float(), size=(H, W), mode="bilinear", align_corners=True)[0]
        top2 = torch.topk(up, 2, dim=0)[0]
        assert (top2[0] - top2[1])[mism].max().item() <= d64 + n32, "argmax differs away from a numerical tie"
    o64, o32, omism, _ = res[False]
    print("[parity] %s (engine's own k-means): |engine-fp64| %.3e  |engine-oracle fp32| %.3e  argmax-equal %.6f"
          % (name, o64, o32, 1.0 - omism.float().mean().item()))
    assert omism.float().mean().item() < 5e-3


def _softmax_properties(probs, preds, first, K):
    seen = set(int(v) for v in torch.unique(first).tolist())
    for p, y in zip(probs, preds):
        assert torch.isfinite(p).all()
        assert p.min().item() >= 0.0 and p.max().item() <= 1.0
        exist = [i for i in range(K + 1) if i in seen]
        s = p[:, exist].sum(1)
        if len(exist) == K + 1:
            assert (s - 1.0).abs().max().item() < 1e-5
        assert set(int(v) for v in torch.unique(y).tolist()) <= seen


def test_720p_k10_properties(model):
    from aocb200.synth import make_clip, restrict_size
    K = 10
    H, W = restrict_size(720, 1280, 10 ** 9)          # native: features 181x321
    frames, labels = make_clip(11, H, W, K, 4)
    first = labels[0].clone()
    first[first == 7] = 0                              # one object id never appears in the ground truth
    a = _run(model, frames, first, K, seed=11, mem_every=2)
    _softmax_properties(a[1], a[0], first, K)
    b = _run(model, frames, first, K, seed=11, mem_every=2)
    c = _run(model, frames, first, K, seed=11, mem_every=2, graphs=False)
    for t in range(3):
        assert torch.equal(a[2][t], b[2][t]), "run-to-run reproducibility (same numpy seed)"
        assert torch.equal(a[2][t], c[2][t]), "graph replay vs plain launches"
        assert torch.equal(a[0][t], c[0][t])
    # Convolution schedules (split-K on / off: different summation orders in the small-map layers), plain launches
    # because a captured graph keeps the schedule it was captured with.  Everything up to the k-means step must agree to
    # fp32 rounding.  The k-means assignment is a discrete decision: when a bank row sits on a cluster boundary a 1e-5
    # change of its embedding flips its label, the centroids move by ~1e-2 and the logits by ~0.1 (measured: 19 of
    # 59 904 rows flip here) -- a property of the reference algorithm (its fp32 and fp64 evaluations differ the same way),
    # so the logits are only held to the tight bound when no label flipped.
    eng = model.engine()
    dumps = []
    for on in (1, 0):
        eng.keep_debug = True
        r = _run(model, frames[:2], first, K, seed=11, mem_every=2, graphs=False, opts=((b"conv_splitk", on),))
        dbg = eng.debug
        dumps.append((r[2][0].clone(), dbg["S"].clone(), dbg["g"].clone(), dbg["loc"].buf.clone(), dbg["labels"].clone()))
        eng.keep_debug = False
    (la, Sa, ga, loca, laba), (lb, Sb, gb, locb, labb) = dumps
    dS, dg, dloc = (Sa - Sb).abs().max().item(), (ga - gb).abs().max().item(), (loca - locb).abs().max().item()
    flips = int((laba != labb).sum().item())
    dl = (la - lb).abs().max().item()
    print("[parity] 720p K=10 split-K vs unsplit convolution schedule: bank embeddings %.3e, global features %.3e, local "
          "features %.3e, k-means label flips %d of %d, logits %.3e" % (dS, dg, dloc, flips, laba.numel(), dl))
    assert dS < 1e-4 and dg < 5e-4 and dloc < 5e-4, (dS, dg, dloc)
    assert dl < (2e-2 if flips == 0 else 1.0), (dl, flips)


def test_1080p_growing_bank_properties(model):
    from aocb200.synth import make_clip, restrict_size
    K = 5
    H, W = restrict_size(1080, 1920, 10 ** 9)         # native: features 269x481, bank +1 frame per step
    frames, labels = make_clip(12, H, W, K, 5)
    a = _run(model, frames, labels[0], K, seed=12, mem_every=1)
    _softmax_properties(a[1], a[0], labels[0], K)
    assert model.engine().bank.n == 4                  # GT frame + frames 1..3 (the last frame is appended by the caller)
    b = _run(model, frames, labels[0], K, seed=12, mem_every=1)
    for t in range(4):
        assert torch.equal(a[2][t], b[2][t])
