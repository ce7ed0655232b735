"""the eval loop's label bookkeeping (label-existence filter, argmax, entropy -> label 125) fused
behind the upsample + softmax kernel, and the device-resident sequence state built on it, against the torch
restatement of eval_manager_mm.py:252-361 (aocb200/sequence.py::run_sequence) on the same inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(state_dict):
    from aocb200.model import get_module
    m = get_module()(None, None)
    m.load_state_dict(state_dict)
    return m.cuda().eval()


@pytest.mark.parametrize("O,h,w,H,W,seen,thr", [
    (6, 31, 54, 120, 214, None, 1.0),             # every slot seen
    (6, 31, 54, 120, 214, [0, 1, 3, 4], 0.8),     # slots 2 and 5 never appeared in a ground-truth frame
    (2, 9, 13, 33, 49, [0, 1], 0.5),
    (16, 17, 19, 65, 77, list(range(0, 16, 2)), 1.2),
    (1, 5, 7, 5, 7, None, 1.0),                   # background only, no resize
])
def test_upsample_softmax_label(model, O, h, w, H, W, seen, thr):
    from aocb200.sequence import shannon_entropy
    eng = model.engine()
    g = torch.Generator().manual_seed(O * 1000 + H)
    logits = (torch.randn(1, O, h, w, generator=g) * 1.5).cuda()
    probs = torch.empty((1, O, H, W), device="cuda")
    label = torch.empty((H, W), dtype=torch.uint8, device="cuda")
    conf = torch.empty((H, W), dtype=torch.uint8, device="cuda")
    ent = torch.empty((H, W), device="cuda")
    eng.set_seen_labels(seen)
    eng.L.upsample_softmax_label_f32(logits.data_ptr(), probs.data_ptr(), label.data_ptr(), conf.data_ptr(),
                                     ent.data_ptr(), eng._exist.data_ptr(), thr, O, h, w, H, W, eng.stream)
    eng.set_seen_labels(None)
    torch.cuda.synchronize()
    # eval_manager_mm.py:252-270 on the reference's own output (aocnet.py:100-107)
    p = torch.softmax(F.interpolate(logits, size=(H, W), mode="bilinear", align_corners=True), dim=1)
    ex = list(range(O)) if seen is None else seen
    keep = torch.zeros(O, device="cuda")
    keep[ex] = 1.0
    assert (probs - p).abs().max().item() < 2e-6    # the probabilities are the reference's plain softmax (aocnet.py:100-107)
    assert (probs.sum(1) - 1.0).abs().max().item() < 1e-5
    p_exist, p = p[:, ex], p * keep.view(1, -1, 1, 1)     # the filter shapes only the derived label maps
    want = torch.argmax(p[0], dim=0)
    mism = label.long() != want
    if mism.any():                                  # only at numerical ties of the two largest probabilities
        top2 = torch.topk(p[0], min(2, O), dim=0).values
        assert O > 1 and ((top2[0] - top2[1])[mism] < 2e-6).all()
    assert mism.float().mean().item() < 1e-3
    u = shannon_entropy(p_exist)[0, 0]
    assert (ent - u).abs().max().item() < 5e-6
    want_c = torch.where(u > thr, torch.full_like(want, 125), want)
    bad = (conf.long() != want_c) & ~mism
    assert ((u - thr).abs()[bad] < 1e-5).all()      # only where the entropy sits on the threshold
    assert bad.float().mean().item() < 1e-3
    if O > 1 and thr < 1.0:
        assert (conf == 125).any() and (conf != 125).any()     # the case exercises both sides of the threshold


@pytest.mark.parametrize("later", [False, True], ids=["first-frame-gt", "object-joins-later"])
def test_device_sequence_vs_eval_loop(model, later):
    """DeviceSequence (fused kernel, uint8 device label maps) against run_sequence (torch ops on the returned
    probabilities): same engine, same RNG stream -> the same label maps, frame by frame.  One object id is absent from
    the first frame (label-existence filter); in the `later` case it enters with ground truth at frame 3."""
    from aocb200.sequence import run_sequence, run_sequence_device
    from aocb200.synth import make_clip
    K, T = 3, 6
    frames, labels = make_clip(33, 65, 97, K, T)
    first = labels[0].clone()
    first[first == 2] = 0
    join = {3: torch.where(labels[3] == 2, labels[3], torch.zeros_like(labels[3]))} if later else None
    dev = torch.device("cuda:0")
    np.random.seed(5)
    want, wprobs = run_sequence(model, frames, first, K, mem_every=2, unc_ratio=0.3, device=dev, keep_probs=True,
                                later_labels=join)
    np.random.seed(5)
    got = run_sequence_device(model, frames, first, K, mem_every=2, unc_ratio=0.3, later_labels=join)
    eng = model.engine()
    assert eng._exist_bits == -1 and eng.unc_ratio == 1.0 and eng.want_probs      # the sequence's settings did not leak
    assert len(got) == len(want) == T - 1
    for t, (a, b) in enumerate(zip(got, want)):
        assert a.dtype == torch.uint8 and a.is_cuda
        eq = (a.long() == b.to(dev)).float().mean().item()
        print("[parity] device sequence frame %d: label agreement %.6f" % (t + 1, eq))
        assert eq == 1.0, (t, eq)
    assert (wprobs[0][:, 2] == 0).all()             # the absent id never wins before it has been seen
    if later:
        assert (got[-1] == 2).any()                 # and is tracked once it has joined


def test_plain_forward_after_device_sequence(model):
    """A DeviceSequence with a label filter must not change what a later plain forward_for_eval caller gets: the
    probabilities stay the reference's softmax over ALL slots (they sum to 1; aocnet.py:100-107)."""
    from aocb200.sequence import run_sequence, run_sequence_device
    from aocb200.synth import make_clip
    K = 3
    frames, labels = make_clip(34, 65, 97, K, 3)
    first = labels[0].clone()
    first[first == 2] = 0
    np.random.seed(6)
    run_sequence_device(model, frames, first, K, mem_every=2, unc_ratio=0.3)
    np.random.seed(6)
    _, probs = run_sequence(model, frames, labels[0], K, mem_every=2, device=torch.device("cuda:0"), keep_probs=True)
    for p in probs:
        assert p.shape[1] == K + 1
        assert (p.sum(1) - 1.0).abs().max().item() < 1e-5
        assert (p[:, 2] > 0).any()


def test_new_sequence_same_shape_one_frame_bank(model):
    """Two DIFFERENT clips of the same size and object count, each with a bank that never grows past the ground-truth
    frame (mem_every = -1): the bank-dependent graph of the first sequence must not be replayed for the second (its
    version key is never reused), so graph replay and plain launches agree bit for bit on both."""
    from aocb200.sequence import run_sequence
    from aocb200.synth import make_clip
    K = 2
    dev = torch.device("cuda:0")
    eng = model.engine()
    clips = [make_clip(s, 65, 97, K, 3) for s in (41, 42)]
    outs = {}
    for graphs in (True, False):
        old, eng.use_graphs = eng.use_graphs, graphs
        try:
            for i, (frames, labels) in enumerate(clips):
                np.random.seed(50 + i)
                outs[(graphs, i)] = run_sequence(model, frames, labels[0], K, mem_every=-1, device=dev)
        finally:
            eng.use_graphs = old
    for i in range(2):
        for a, b in zip(outs[(True, i)], outs[(False, i)]):
            assert torch.equal(a, b), "graph replay differs from plain launches on sequence %d" % i


def test_object_count_changes_between_sequences(model):
    """K is read from the gt_ids of THIS call (fresh tensors per sequence may recycle Python ids)."""
    from aocb200.sequence import run_sequence
    from aocb200.synth import make_clip
    dev = torch.device("cuda:0")
    for K in (3, 1, 2):
        frames, labels = make_clip(60 + K, 65, 97, K, 3)
        np.random.seed(K)
        _, probs = run_sequence(model, frames, labels[0], K, mem_every=2, device=dev, keep_probs=True)
        assert all(p.shape[1] == K + 1 for p in probs)
