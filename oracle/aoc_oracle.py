"""CPU oracle for the AOC-Net per-frame inference path (TEST INFRASTRUCTURE ONLY).

This file is a from-scratch fp32 restatement of the reference's `forward_for_eval` hot path in
plain torch-on-CPU + numpy/scipy.  It is the checker for the CUDA engine in ``aocb200/``; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import it.  The product path never does.

Parity status: **pinned against the reference itself** -- tools/make_golden.py imports the
(repaired) reference from /root/reference in the build container, runs both on
identical seeded inputs/weights/k-means RNG stream, asserts agreement, and stores the reference's
outputs as fixtures in tests/golden/.  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4).  Third-party arithmetic on the path: ``scipy.cluster.vq.kmeans2`` (unpinned in the
reference's Dockerfile:69; this image has scipy 1.18) -- called here exactly as the reference
does, and restated in numpy in ``kmeans2_points`` below for documentation and for the CUDA
kernel's parity tests.

All paths below are relative to /root/reference/AOC-Net/complete_project/AOCNet/.
Weights come as a state_dict with the reference's parameter names.
"""
import numpy as np
import torch
import torch.nn.functional as F

WRONG_LABEL_PADDING_DISTANCE = 5e4  # networks/layers/matching.py:25


# --------------------------------------------------------------------------------------
# k-means (third-party: scipy.cluster.vq.kmeans2, minit='points')
# --------------------------------------------------------------------------------------
def kmeans2_points(data, k, iters=20, rng=None):
    """numpy restatement of scipy.cluster.vq.kmeans2(data, k, minit='points', iter=iters).

    scipy/cluster/vq/_vq_impl.py: init = data[rng.choice(n, k, replace=False)] (_kpoints :465),
    then `iters` rounds of {assign to nearest code (lowest index wins ties), mean per code,
    empty codes keep their previous position} (:773-783).  The returned labels are those of the
    last round, i.e. computed against the code book *before* its final update.
    """
    rng = np.random.mtrand._rand if rng is None else rng
    data = np.asarray(data)
    n = data.shape[0]
    idx = rng.choice(n, size=int(k), replace=False)
    code = data[idx].copy()
    label = np.zeros(n, dtype=np.int32)
    x2 = (data * data).sum(1)
    for _ in range(iters):
        c2 = (code * code).sum(1)
        d = (data @ code.T) * data.dtype.type(-2.0) + x2[:, None] + c2[None, :]
        label = np.argmin(d, axis=1).astype(np.int32)
        new = np.zeros_like(code)
        cnt = np.bincount(label, minlength=int(k))
        np.add.at(new, label, data)
        has = cnt > 0
        new[has] /= cnt[has, None].astype(data.dtype)
        new[~has] = code[~has]
        code = new
    return code, label


def _scipy_kmeans2(x_np, k):
    from scipy.cluster.vq import kmeans2
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return kmeans2(x_np, k, minit="points", iter=20)


# --------------------------------------------------------------------------------------
# small layer helpers
# --------------------------------------------------------------------------------------
class _W:
    """Prefix view on a state_dict."""

    def __init__(self, sd, prefix=""):
        self.sd, self.p = sd, prefix

    def __call__(self, name):
        return self.sd[self.p + name]

    def has(self, name):
        return (self.p + name) in self.sd

    def sub(self, name):
        return _W(self.sd, self.p + name + ".")


def frozen_bn(x, w):
    # networks/layers/normalization.py:17-23
    scale = w("weight") * (w("running_var") + 1e-5).rsqrt()
    bias = w("bias") - w("running_mean") * scale
    return x * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)


def conv(x, w, stride=1, padding=0, dilation=1, groups=1):
    b = w("bias") if w.has("bias") else None
    return F.conv2d(x, w("weight"), b, stride, padding, dilation, groups)


def group_norm(x, w, groups):
    return F.group_norm(x, groups, w("weight"), w("bias"), 1e-5)


def linear(x, w):
    return F.linear(x, w("weight"), w("bias"))


def gct(x, w, eps=1e-5):
    # networks/layers/gct.py:17-36 (mode 'l2')
    emb = (x.pow(2).sum((2, 3), keepdim=True) + eps).pow(0.5) * w("alpha")
    norm = w("gamma") / (emb.pow(2).mean(dim=1, keepdim=True) + eps).pow(0.5)
    return x * (1.0 + torch.tanh(emb * norm + w("beta")))


def ia_gate(x, head, w):
    # networks/layers/attention.py:12-17
    a = 1.0 + torch.tanh(linear(head, w.sub("IA")))
    return a[:, :, None, None] * x


def gn_bottleneck(x, w, stride=1, dilation=1):
    # networks/layers/gct.py:68-91 (GroupNorm(32) bottleneck with a GCT gate in front)
    out = gct(x, w.sub("GCT1"))
    out = F.relu(group_norm(conv(out, w.sub("conv1")), w.sub("bn1"), 32))
    out = conv(out, w.sub("conv2"), stride=stride, padding=dilation, dilation=dilation)
    out = F.relu(group_norm(out, w.sub("bn2"), 32))
    out = group_norm(conv(out, w.sub("conv3")), w.sub("bn3"), 32)
    if w.has("downsample.0.weight"):
        res = group_norm(conv(x, w.sub("downsample.0"), stride=stride), w.sub("downsample.1"), 32)
    else:
        res = x
    return F.relu(out + res)


# --------------------------------------------------------------------------------------
# backbone (networks/deeplab/*)
# --------------------------------------------------------------------------------------
def _res_block(x, w, stride, dilation):
    # networks/deeplab/backbone/resnet.py:23-42
    out = F.relu(frozen_bn(conv(x, w.sub("conv1")), w.sub("bn1")))
    out = conv(out, w.sub("conv2"), stride=stride, padding=dilation, dilation=dilation)
    out = F.relu(frozen_bn(out, w.sub("bn2")))
    out = frozen_bn(conv(out, w.sub("conv3")), w.sub("bn3"))
    if w.has("downsample.0.weight"):
        x = frozen_bn(conv(x, w.sub("downsample.0"), stride=stride), w.sub("downsample.1"))
    return F.relu(out + x)


RESNET101_LAYERS = [  # (name, blocks, stride of first block, dilation per block)  resnet.py:49-65,143-149
    ("layer1", 3, 1, [1, 1, 1]),
    ("layer2", 4, 2, [1] * 4),
    ("layer3", 23, 2, [1] * 23),
    ("layer4", 3, 1, [2, 4, 8]),
]


def resnet101(x, w):
    # networks/deeplab/backbone/resnet.py:108-123
    x = F.relu(frozen_bn(conv(x, w.sub("conv1"), stride=2, padding=3), w.sub("bn1")))
    x = F.max_pool2d(x, 3, 2, 1)
    low = None
    for name, blocks, stride, dils in RESNET101_LAYERS:
        for b in range(blocks):
            x = _res_block(x, w.sub("%s.%d" % (name, b)), stride if b == 0 else 1, dils[b])
        if name == "layer1":
            low = x
    return x, low


def deeplab_aspp(x, w):
    # networks/deeplab/aspp.py:62-74 (dropout is identity in eval)
    outs = [F.relu(frozen_bn(conv(x, w.sub("aspp1.atrous_conv")), w.sub("aspp1.bn")))]
    for i, d in ((2, 6), (3, 12), (4, 18)):
        p = w.sub("aspp%d" % i)
        outs.append(F.relu(frozen_bn(conv(x, p.sub("atrous_conv"), padding=d, dilation=d), p.sub("bn"))))
    g = x.mean((2, 3), keepdim=True)
    g = F.relu(frozen_bn(conv(g, w.sub("global_avg_pool.1")), w.sub("global_avg_pool.2")))
    outs.append(F.interpolate(g, size=x.shape[2:], mode="bilinear", align_corners=True))
    x = torch.cat(outs, 1)
    return F.relu(frozen_bn(conv(x, w.sub("conv1")), w.sub("bn1")))


def deeplab_decoder(x, low, w):
    # networks/deeplab/decoder.py:32-41
    low = F.relu(frozen_bn(conv(low, w.sub("conv1")), w.sub("bn1")))
    x = F.interpolate(x, size=low.shape[2:], mode="bilinear", align_corners=True)
    x = torch.cat((x, low), 1)
    x = F.relu(frozen_bn(conv(x, w.sub("last_conv.0"), padding=1), w.sub("last_conv.1")))
    x = F.relu(frozen_bn(conv(x, w.sub("last_conv.4"), padding=1), w.sub("last_conv.5")))
    return x


def extract_feature(img, sd):
    """networks/aoc/aocnet.py:109-112 + deeplab/deeplab.py:27-38.  -> (emb[1,100,h,w], low[1,256,h,w])"""
    w = _W(sd)
    fe = w.sub("feature_extracter")
    x, low = resnet101(img, fe.sub("backbone"))
    x = deeplab_decoder(deeplab_aspp(x, fe.sub("aspp")), low, fe.sub("decoder"))
    # semantic_embedding: aocnet.py:19-25
    x = conv(x, w.sub("seperate_conv"), padding=1, groups=x.shape[1])
    x = F.relu(group_norm(x, w.sub("bn1"), 32))
    x = conv(x, w.sub("embedding_conv"))
    x = F.relu(group_norm(x, w.sub("bn2"), 25))
    return x, low


# --------------------------------------------------------------------------------------
# matching (networks/layers/matching.py)
# --------------------------------------------------------------------------------------
def _sig(d, bias):
    return (torch.sigmoid(d + bias) - 0.5) * 2


def _pairwise(q, q2, r, r2):
    # matching.py:41-46: (|q|^2 + |r|^2) - 2 q.r
    return q2.unsqueeze(1) + r2.unsqueeze(0) - 2.0 * torch.matmul(q, r.t())


def compact_bank(ref_embs, ref_onehots):
    """matching.py:2486-2492: concatenate frames (frame-major raster), drop pixels of no object.
    ref_embs: list of [h,w,C]; ref_onehots: list of [h,w,O].  -> (emb[N,C], onehot[N,O]) or None."""
    C = ref_embs[0].shape[-1]
    O = ref_onehots[0].shape[-1]
    e = torch.cat([x.reshape(-1, C) for x in ref_embs], 0)
    l = torch.cat([x.reshape(-1, O) for x in ref_onehots], 0)
    keep = l.sum(1) > 0.9
    if int(keep.sum()) == 0:
        return None
    return e[keep], l[keep]


def global_matching_for_eval(ref_embs, q, ref_onehots, dis_bias):
    """matching.py:2384-2510 -> [h,w,O] (the trailing feature dim of 1 is dropped)."""
    h, w, C = q.shape
    O = ref_onehots[0].shape[-1]
    bank = compact_bank(ref_embs, ref_onehots)
    if bank is None:
        return torch.ones(h, w, O)
    e, l = bank
    qf = q.reshape(-1, C)
    q2, e2 = qf.pow(2).sum(1), e.pow(2).sum(1)
    # matching.py:200-249 walks the query pixels in chunks (rows are independent); here the chunk only bounds the
    # [chunk, N] distance block to ~1 Gi elements -- the golden-fixture sizes are a single chunk
    step = max(1, min(qf.shape[0], (1 << 30) // max(1, e.shape[0])))
    rows = []
    for c0 in range(0, qf.shape[0], step):
        d = _pairwise(qf[c0:c0 + step], q2[c0:c0 + step], e, e2)  # [chunk, N]
        feats = []
        for o in range(O):
            # matching.py:83-90: min_n (d + 5e4*[label_n != o]); fp add is monotone so the min splits.
            right = l[:, o] > 0.1
            cand = []
            if bool(right.any()):
                cand.append(d[:, right].min(1)[0])
            if bool((~right).any()):
                cand.append(d[:, ~right].min(1)[0] + WRONG_LABEL_PADDING_DISTANCE)
            feats.append(cand[0] if len(cand) == 1 else torch.minimum(cand[0], cand[1]))
        rows.append(torch.stack(feats, 1))
        del d
    feats = torch.cat(rows, 0).view(h, w, O)
    return _sig(feats, dis_bias.view(1, 1, -1))


def adaptive_proxies(e, l, cluster_num=16, kmeans_fn=None):
    """matching.py:533-595: per-object k-means proxies.  Returns list over objects of
    None | (centroid[k,C], centroid_avg[k',C])."""
    kmeans_fn = kmeans_fn or _scipy_kmeans2
    out = []
    for o in range(l.shape[1]):
        idx = torch.nonzero(l[:, o] > 0.9).squeeze(1)
        x = e.index_select(0, idx)
        cluster_num = min(cluster_num, x.shape[0])  # :556 -- carries over to later objects
        if cluster_num == 0:
            out.append(None)
            continue
        try:
            cen, lab = kmeans_fn(x.numpy(), cluster_num)
        except Exception:
            out.append(None)  # :592-595
            continue
        lab_t = torch.from_numpy(np.asarray(lab)).long()
        # :589 -- rows of the ALL-object bank at object-local indices, non-empty labels only.
        avg = torch.stack([e.index_select(0, torch.nonzero(lab_t == j).squeeze(1)).sum(0) / float((lab == j).sum())
                           for j in np.unique(lab)], 0)
        out.append((torch.from_numpy(np.asarray(cen)), avg))
    return out


def global_matching_for_eval_cluster(ref_embs, q, ref_onehots, dis_bias, kmeans_fn=None):
    """matching.py:1571-1705 + :506-640 -> [h,w,O,2]."""
    h, w, C = q.shape
    O = ref_onehots[0].shape[-1]
    bank = compact_bank(ref_embs, ref_onehots)
    if bank is None:
        return torch.ones(h, w, O, 1)  # :1679-1680 (latent shape quirk, last dim 1)
    e, l = bank
    qf = q.reshape(-1, C)
    q2 = qf.pow(2).sum(1)
    prox = adaptive_proxies(e, l, 16, kmeans_fn)
    maps = [[], []]
    for o in range(O):
        for s in range(2):
            if prox[o] is None:
                maps[s].append(torch.full((qf.shape[0],), WRONG_LABEL_PADDING_DISTANCE))
            else:
                c = prox[o][s]
                maps[s].append(_pairwise(qf, q2, c, c.pow(2).sum(1)).min(1)[0])
    outs = [_sig(torch.stack(m, 1).view(h, w, O), dis_bias.view(1, 1, -1)) for m in maps]
    return torch.stack(outs, -1)


def global_matching_for_eval_proxy(proxies, q, dis_bias):
    """matching.py:2518-2662 (+:149-197): distance to each object's single mean proxy -> [h,w,O]."""
    h, w, C = q.shape
    qf = q.reshape(-1, C)
    d = _pairwise(qf, qf.pow(2).sum(1), proxies, proxies.pow(2).sum(1))
    return _sig(d.view(h, w, -1), dis_bias.view(1, 1, -1))


def local_pairwise_distances(x, y, max_distance=12):
    """matching.py:2710-2755 -> d[hh,ww,(2m+1)^2] on the half-resolution grid."""
    H, W, C = x.shape
    x = x.permute(2, 0, 1).unsqueeze(0)
    y = y.permute(2, 0, 1).unsqueeze(0)
    size = (H // 2 + 1, W // 2 + 1)
    x = F.interpolate(x, size=size, mode="bilinear", align_corners=True)
    y = F.interpolate(y, size=size, mode="bilinear", align_corners=True)
    hh, ww = size
    m = max_distance
    x2 = x.pow(2).sum(1).view(hh, ww, 1)
    y2 = y.pow(2).sum(1).view(1, 1, hh, ww)
    py = F.pad(y, (m, m, m, m))
    py2 = F.pad(y2, (m, m, m, m), value=WRONG_LABEL_PADDING_DISTANCE)
    oy = F.unfold(py, kernel_size=(hh, ww)).view(C, hh * ww, -1).permute(1, 0, 2)
    oy2 = F.unfold(py2, kernel_size=(hh, ww)).view(hh, ww, -1)
    xf = x.view(C, hh * ww, 1).permute(1, 2, 0)
    return x2 + oy2 - 2.0 * torch.matmul(xf, oy).view(hh, ww, -1)


def local_matching(prev_emb, q, prev_onehot, dis_bias, dists=(2, 4, 6, 8, 10, 12)):
    """matching.py:2757-2851 (local_matching_proxy :2853-2947 has the same body) -> [h,w,O,6].
    Channel order: window 12 first, then 2,4,6,8,10."""
    H, W, _ = q.shape
    O = prev_onehot.shape[-1]
    m = dists[-1]
    d = local_pairwise_distances(q, prev_emb, m)
    hh, ww = d.shape[:2]
    lab = prev_onehot.permute(2, 0, 1).unsqueeze(1)
    if (hh, ww) != (H, W):
        lab = F.interpolate(lab, size=(hh, ww), mode="nearest")
    plab = F.pad(lab, (m, m, m, m), value=0)
    masks = F.unfold(plab, kernel_size=(hh, ww)).view(O, hh, ww, -1).permute(1, 2, 3, 0) > 0.9
    pad = torch.tensor(WRONG_LABEL_PADDING_DISTANCE)
    dm = torch.where(masks, d.unsqueeze(-1).expand(-1, -1, -1, O), pad)
    outs = [dm.min(2)[0]]
    dm5 = dm.view(hh, ww, 2 * m + 1, 2 * m + 1, O)
    for r in dists[:-1]:
        sub = dm5[:, :, m - r:m + r + 1, m - r:m + r + 1, :].reshape(hh, ww, -1, O)
        outs.append(sub.min(2)[0])
    md = torch.stack(outs, 0).permute(3, 0, 1, 2)  # [O,6,hh,ww]
    md = _sig(md, dis_bias.view(-1, 1, 1, 1))
    if (hh, ww) != (H, W):
        md = F.interpolate(md, size=(H, W), mode="bilinear", align_corners=True)
    return md.permute(2, 3, 0, 1)


def foreground2background(dis):
    """matching.py:9-23: per object, min over the OTHER objects.  dis: [O, ...]."""
    O = dis.shape[0]
    if O == 1:
        return dis
    return torch.stack([torch.cat([dis[:i], dis[i + 1:]], 0).min(0)[0] for i in range(O)], 0)


def attention_heads(ref_embs, ref_onehots, prev_emb, prev_onehot, eps=1e-5):
    """networks/layers/attention.py:155-189.  ref_embs: list [1,C,h,w]; ref_onehots: list [O,1,h,w];
    prev_emb [1,C,h,w]; prev_onehot [O,1,h,w] -> (head[O,4C], ref_pos, ref_neg, prev_pos, prev_neg)."""
    tp = tn = np_ = nn_ = 0.0
    for e, l in zip(ref_embs, ref_onehots):
        pos = (e * l).sum((2, 3))
        tp = tp + pos
        tn = tn + (e.sum((2, 3)) - pos)
        np_ = np_ + l.sum((2, 3))
        nn_ = nn_ + (1.0 - l).sum((2, 3))
    ref_pos = tp / (np_ + eps)
    ref_neg = tn / (nn_ + eps)
    pos = (prev_emb * prev_onehot).sum((2, 3))
    neg = prev_emb.sum((2, 3)) - pos
    prev_pos = pos / (prev_onehot.sum((2, 3)) + eps)
    prev_neg = neg / ((1.0 - prev_onehot).sum((2, 3)) + eps)
    return torch.cat([ref_pos, ref_neg, prev_pos, prev_neg], 1), ref_pos, ref_neg, prev_pos, prev_neg


# --------------------------------------------------------------------------------------
# calibration decoder (networks/aoc/decoding_module.py, conditioning_layer.py, layers/aspp.py)
# --------------------------------------------------------------------------------------
def conditioning_layer(z, w, beta):
    """networks/aoc/conditioning_layer.py:24-48 with repairs R5/R7.  z [O,C,h,w] -> [O,C]."""
    phi = conv(z, w.sub("phi_layer")).reshape(z.shape[0], 1, -1)
    zf = z.reshape(z.shape[0], z.shape[1], -1)
    rank = max(1, int(beta * z.shape[-1] * z.shape[-2]))
    kth = torch.topk(phi, k=rank, dim=-1, sorted=True)[0][..., -1, None]
    gap = (zf * (phi > kth)).mean(-1)  # strict '>' and mean over ALL positions
    return linear(gap, w.sub("mlp_layer"))


def conditioning_block(x, head, w, beta=0.3):
    """conditioning_layer.py:63-86 with repairs R4/R6."""
    px = x.mean((2, 3), keepdim=True)
    delta = px.sum(0, keepdim=True) - px
    c1 = conditioning_layer(x, w.sub("CL_1"), beta)
    c2 = conditioning_layer(delta, w.sub("CL_2"), beta)
    c3 = conditioning_layer(head[:, :, None, None], w.sub("CL_3"), 1)
    a = 1.0 + torch.tanh(linear(torch.cat([c1, c2, c3], 1), w.sub("mlp_layer")))
    return a[:, :, None, None] * x


def decoder_aspp(x, w):
    # networks/layers/aspp.py:57-70
    outs = []
    for i, d in ((1, 0), (2, 6), (3, 12), (4, 18)):
        p = w.sub("aspp%d" % i)
        y = conv(gct(x, p.sub("GCT")), p.sub("atrous_conv"), padding=d, dilation=max(d, 1))
        outs.append(F.relu(group_norm(y, p.sub("bn"), 32)))
    g = F.relu(conv(x.mean((2, 3), keepdim=True), w.sub("global_avg_pool.1")))
    outs.append(F.interpolate(g, size=x.shape[2:], mode="bilinear", align_corners=True))
    x = gct(torch.cat(outs, 1), w.sub("GCT"))
    return F.relu(group_norm(conv(x, w.sub("conv1")), w.sub("bn1"), 32))


def _delta_head(x, head):
    px = x.mean((2, 3))
    return torch.cat([head, px.sum(0, keepdim=True) - px], 1)


def _modulator(x, mem, head, w, tag):
    # decoding_module.py:192-210
    x = torch.cat([x, mem], 1)
    for i in (1, 2, 3):
        x = ia_gate(x, head, w.sub("%s_Reweight_Layer_%d" % (tag, i)))
        x = gn_bottleneck(x, w.sub("%s_Bottleneck_%d" % (tag, i)))
    return x


def _dyn_logit(x, head, w):
    # decoding_module.py:151-160: per-object 1x1 conv whose weights come from the head.
    o = linear(head, w)
    c = x.shape[1]
    return (x * o[:, :c, None, None]).sum(1, keepdim=True) + o[:, -1].view(-1, 1, 1, 1)


def calibration_decoding(x, head, memory, low, sd_prefix_w):
    """decoding_module.py:96-149 (repairs R2/R3/R8/R9).  x [O,164,h,w]; head [O,400];
    memory [m0|None, m1|None]; low [1,256,h,w] -> (logits [1,O,h,w], [m0, m1])."""
    w = sd_prefix_w
    x = ia_gate(x, head, w.sub("IA1"))
    x = gn_bottleneck(x, w.sub("layer1"))
    x = conditioning_block(x, head, w.sub("CLB2"))
    x = gn_bottleneck(x, w.sub("layer2"), 1, 2)
    x = conditioning_block(x, head, w.sub("CLB3"))
    x = gn_bottleneck(x, w.sub("layer3"), 2, 1)
    x = conditioning_block(x, head, w.sub("CLB4"))
    x = gn_bottleneck(x, w.sub("layer4"), 1, 2)
    x = conditioning_block(x, head, w.sub("CLB5"))
    x = gn_bottleneck(x, w.sub("layer5"), 1, 4)
    x = ia_gate(x, _delta_head(x, head), w.sub("IA9"))
    x = decoder_aspp(x, w.sub("ASPP"))
    cur1 = x
    m0 = memory[0] if (memory[0] is not None and memory[0].shape == cur1.shape) else cur1
    x = _modulator(x, m0, head, w, "M1")
    cur2 = x
    m1 = memory[1] if (memory[1] is not None and memory[1].shape == cur2.shape) else cur2
    x = _modulator(x, m1, head, w, "M2")
    # decoder_final :162-190
    x = F.interpolate(x, size=low.shape[2:], mode="bicubic", align_corners=True)
    sc = gct(torch.cat([low.expand(x.shape[0], -1, -1, -1), x], 1), w.sub("GCT_sc"))  # R9
    sc = F.relu(group_norm(conv(sc, w.sub("conv_sc")), w.sub("bn_sc"), 16))
    x = torch.cat([x, sc], 1)
    x = ia_gate(x, _delta_head(x, head), w.sub("IA10"))
    x = F.relu(group_norm(conv(x, w.sub("conv1"), padding=1), w.sub("bn1"), 32))
    x = ia_gate(x, _delta_head(x, head), w.sub("IA11"))
    x = F.relu(group_norm(conv(x, w.sub("conv2"), padding=1), w.sub("bn2"), 32))
    fg = _dyn_logit(x, head, w.sub("IA_final_fg"))
    bg = _dyn_logit(x, head, w.sub("IA_final_bg"))
    # augment_background_logit :213-225
    pred = fg.clone()
    if fg.shape[0] > 1:
        pred[0:1] = pred[0:1] + bg[1:].min(0, keepdim=True)[0]
    return pred.permute(1, 0, 2, 3), [cur1, m1]


# --------------------------------------------------------------------------------------
# the per-frame entry (networks/aoc/aocnet.py:84-372)
# --------------------------------------------------------------------------------------
class AOCOracle:
    """Same call contract as the reference's AOCNet.forward_for_eval (aocnet.py:84-107), bs == 1."""

    def __init__(self, state_dict, kmeans_fn=None):
        self.sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
        self.kmeans_fn = kmeans_fn
        self.last_logits = None
        self.last_prehead_in = None

    def extract_feature(self, img):
        return extract_feature(img.float(), self.sd)

    def match_features(self, ref_embs, ref_masks, prev_emb, prev_mask, cur_emb, K):
        """aocnet.py:128-358: returns (pre_to_cat [O,24,h,w], head [O,400], prev_onehot [O,1,h,w])."""
        sd = self.sd
        h, w = cur_emb.shape[2:]
        ids = torch.arange(0, K + 1).int().view(-1, 1, 1, 1)
        O = K + 1
        bias = torch.cat([sd["bg_bias"].view(1), sd["fg_bias"].view(1).expand(K)]) if K > 0 else sd["bg_bias"].view(1)

        def onehot(mask):
            s = F.interpolate(mask.float(), size=(h, w), mode="nearest").int()
            return (s[0] == ids).float()  # [O,1,h,w]

        ref_oh = [onehot(m) for m in ref_masks]
        prev_oh = onehot(prev_mask)
        q = cur_emb[0].permute(1, 2, 0)
        ref_hw = [e[0].permute(1, 2, 0) for e in ref_embs]
        ref_oh_hw = [l.squeeze(1).permute(1, 2, 0) for l in ref_oh]
        prev_oh_hw = prev_oh.squeeze(1).permute(1, 2, 0)

        g = global_matching_for_eval(ref_hw, q, ref_oh_hw, bias)                    # [h,w,O]
        gc = global_matching_for_eval_cluster(ref_hw, q, ref_oh_hw, bias, self.kmeans_fn)  # [h,w,O,2]
        loc = local_matching(prev_emb[0].permute(1, 2, 0), q, prev_oh_hw, bias)     # [h,w,O,6]
        head, ref_pos, _, prev_pos, _ = attention_heads(ref_embs, ref_oh, prev_emb, prev_oh)
        gp = global_matching_for_eval_proxy(ref_pos, q, bias)                       # [h,w,O]
        inst = torch.matmul(prev_oh_hw, prev_pos)                                   # aocnet.py:325
        locp = local_matching(inst, q, prev_oh_hw, bias)                            # [h,w,O,6]

        g_c = g.permute(2, 0, 1).unsqueeze(1)              # [O,1,h,w]
        gc_c = gc.permute(2, 3, 0, 1)                      # [O,2,h,w]
        gp_c = gp.permute(2, 0, 1).unsqueeze(1)
        loc_c = loc.permute(2, 3, 0, 1)                    # [O,6,h,w]
        locp_c = locp.permute(2, 3, 0, 1)
        g_bg = foreground2background(g_c)
        loc_bg = foreground2background(loc_c)
        pre = torch.cat([g_c, gc_c, gp_c, loc_c, locp_c, prev_oh, loc_bg, g_bg], 1)  # aocnet.py:355-358
        return pre, head, prev_oh

    def forward_for_eval(self, memory_prev_list, ref_embeddings, ref_masks, prev_embedding, prev_mask,
                         current_frame, pred_size, gt_ids):
        sd = self.sd
        emb, low = self.extract_feature(current_frame)
        if prev_embedding is None:
            return None, emb, memory_prev_list
        K = int(gt_ids[0])
        pre, head, _ = self.match_features(ref_embeddings, ref_masks, prev_embedding, prev_mask, emb, K)
        self.last_prehead_in = pre
        w = _W(sd)
        ph = w.sub("dynamic_prehead")
        pre = F.relu(group_norm(conv(pre, ph.sub("conv")), ph.sub("bn"), 16))  # decoding_module.py:236-240
        x = torch.cat([emb.expand(K + 1, -1, -1, -1), pre], 1)
        logits, mem = calibration_decoding(x, head, list(memory_prev_list[0]), low, w.sub("dynamic_seghead"))
        self.last_logits = logits
        pred = F.interpolate(logits, size=(int(pred_size[0]), int(pred_size[1])), mode="bilinear",
                             align_corners=True)
        return torch.softmax(pred, dim=1), emb, [mem]
