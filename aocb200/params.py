"""Parameter inventory of AOC-Net with the reference's state_dict names, and a synthetic checkpoint.

The reference keeps its weights in an nn.Module tree (networks/aoc/aocnet.py:12-51,
networks/aoc/decoding_module.py:11-94, networks/deeplab/*).  The engine only needs the tensors, but
`load_network` (utils/checkpoint.py:49-70) and the eval loop need `.state_dict()/.load_state_dict()`
with the same names, so `ParamTree` rebuilds the name tree from a flat spec.

No pretrained checkpoint exists offline (README.md:93-95 links only), so `synthetic_state_dict`
makes a seeded, calibrated one: every path is exercised (non-trivial GCT
gates, biases, BN statistics) and activations stay O(1).
"""
from collections import OrderedDict

import torch
import torch.nn as nn

EMB = 100          # cfg.MODEL_SEMANTIC_EMBEDDING_DIM
HEAD = 4 * EMB     # attention head width (aocnet.py:39)
PRE = 64           # cfg.MODEL_PRE_HEAD_EMBEDDING_DIM
DEC = 256          # cfg.MODEL_HEAD_EMBEDDING_DIM
REFINE = 64        # cfg.MODEL_REFINE_CHANNELS
PREHEAD_IN = 24    # aocnet.py:43-46


def _bn(spec, p, c):
    for n in ("weight", "bias", "running_mean", "running_var"):
        spec[p + "." + n] = ("bn_" + n, (c,))


def _gn(spec, p, c):
    spec[p + ".weight"] = ("gn_weight", (c,))
    spec[p + ".bias"] = ("gn_bias", (c,))


def _conv(spec, p, cout, cin, k, bias=False):
    spec[p + ".weight"] = ("conv", (cout, cin, k, k))
    if bias:
        spec[p + ".bias"] = ("bias", (cout,))


def _lin(spec, p, cout, cin):
    spec[p + ".weight"] = ("linear", (cout, cin))
    spec[p + ".bias"] = ("bias", (cout,))


def _gct(spec, p, c):
    spec[p + ".alpha"] = ("gct_alpha", (1, c, 1, 1))
    spec[p + ".gamma"] = ("gct_gamma", (1, c, 1, 1))
    spec[p + ".beta"] = ("gct_beta", (1, c, 1, 1))


def _res_bottleneck(spec, p, cin, planes, down):
    _conv(spec, p + ".conv1", planes, cin, 1); _bn(spec, p + ".bn1", planes)
    _conv(spec, p + ".conv2", planes, planes, 3); _bn(spec, p + ".bn2", planes)
    _conv(spec, p + ".conv3", planes * 4, planes, 1); _bn(spec, p + ".bn3", planes * 4)
    if down:
        _conv(spec, p + ".downsample.0", planes * 4, cin, 1); _bn(spec, p + ".downsample.1", planes * 4)


def _gn_bottleneck(spec, p, cin, cout):
    planes = cout // 4
    _gct(spec, p + ".GCT1", cin)
    _conv(spec, p + ".conv1", planes, cin, 1); _gn(spec, p + ".bn1", planes)
    _conv(spec, p + ".conv2", planes, planes, 3); _gn(spec, p + ".bn2", planes)
    _conv(spec, p + ".conv3", cout, planes, 1); _gn(spec, p + ".bn3", cout)
    return p


def _cond_block(spec, p, c):
    for name, d in (("CL_1", c), ("CL_2", c), ("CL_3", HEAD)):
        _conv(spec, "%s.%s.phi_layer" % (p, name), 1, d, 1, bias=True)
        _lin(spec, "%s.%s.mlp_layer" % (p, name), d, d)
    _lin(spec, p + ".mlp_layer", c, 2 * c + HEAD)


def param_spec():
    """Ordered {name: (kind, shape)} in the reference's registration order."""
    s = OrderedDict()
    s["bg_bias"] = ("dis_bias", (1, 1, 1, 1))
    s["fg_bias"] = ("dis_bias", (1, 1, 1, 1))
    b = "feature_extracter.backbone"
    _conv(s, b + ".conv1", 64, 3, 7); _bn(s, b + ".bn1", 64)
    cin = 64
    for name, planes, blocks in (("layer1", 64, 3), ("layer2", 128, 4), ("layer3", 256, 23), ("layer4", 512, 3)):
        for i in range(blocks):
            _res_bottleneck(s, "%s.%s.%d" % (b, name, i), cin, planes, i == 0)
            cin = planes * 4
    a = "feature_extracter.aspp"
    for i, k in ((1, 1), (2, 3), (3, 3), (4, 3)):
        _conv(s, "%s.aspp%d.atrous_conv" % (a, i), 256, 2048, k); _bn(s, "%s.aspp%d.bn" % (a, i), 256)
    _conv(s, a + ".global_avg_pool.1", 256, 2048, 1); _bn(s, a + ".global_avg_pool.2", 256)
    _conv(s, a + ".conv1", 256, 1280, 1); _bn(s, a + ".bn1", 256)
    d = "feature_extracter.decoder"
    _conv(s, d + ".conv1", 48, 256, 1); _bn(s, d + ".bn1", 48)
    _conv(s, d + ".last_conv.0", 256, 304, 3); _bn(s, d + ".last_conv.1", 256)
    _conv(s, d + ".last_conv.4", 256, 256, 3); _bn(s, d + ".last_conv.5", 256)
    # semantic embedding is registered twice (attribute + nn.Sequential alias): aocnet.py:19-25
    for names in (("seperate_conv", "bn1", "embedding_conv", "bn2"),
                  ("semantic_embedding.0", "semantic_embedding.1", "semantic_embedding.3", "semantic_embedding.4")):
        s[names[0] + ".weight"] = ("conv", (256, 1, 3, 3)); s[names[0] + ".bias"] = ("bias", (256,))
        _gn(s, names[1], 256)
        _conv(s, names[2], EMB, 256, 1, bias=True)
        _gn(s, names[3], EMB)
    h = "dynamic_seghead"
    cin0 = EMB + PRE
    _lin(s, h + ".IA1.IA", cin0, HEAD)
    _gn_bottleneck(s, h + ".layer1", cin0, DEC)
    _conv(s, h + ".layer1.downsample.0", DEC, cin0, 1); _gn(s, h + ".layer1.downsample.1", DEC)
    _gn_bottleneck(s, h + ".layer2", DEC, DEC)
    _cond_block(s, h + ".CLB2", DEC)
    _gn_bottleneck(s, h + ".layer3", DEC, DEC * 2)
    _conv(s, h + ".layer3.downsample.0", DEC * 2, DEC, 1); _gn(s, h + ".layer3.downsample.1", DEC * 2)
    _cond_block(s, h + ".CLB3", DEC)
    _gn_bottleneck(s, h + ".layer4", DEC * 2, DEC * 2)
    _cond_block(s, h + ".CLB4", DEC * 2)
    _gn_bottleneck(s, h + ".layer5", DEC * 2, DEC * 2)
    _cond_block(s, h + ".CLB5", DEC * 2)
    _lin(s, h + ".IA9.IA", DEC * 2, HEAD + DEC * 2)
    for i, k in ((1, 1), (2, 3), (3, 3), (4, 3)):
        p = "%s.ASPP.aspp%d" % (h, i)
        _gct(s, p + ".GCT", 512); _conv(s, p + ".atrous_conv", 128, 512, k); _gn(s, p + ".bn", 128)
    _conv(s, h + ".ASPP.global_avg_pool.1", 128, 512, 1)
    _gct(s, h + ".ASPP.GCT", 640); _conv(s, h + ".ASPP.conv1", 256, 640, 1); _gn(s, h + ".ASPP.bn1", 256)
    for tag in ("M1", "M2"):
        for i, (ci, co) in enumerate(((512, 512), (512, 256), (256, 256)), 1):
            _lin(s, "%s.%s_Reweight_Layer_%d.IA" % (h, tag, i), ci, HEAD)
            p = "%s.%s_Bottleneck_%d" % (h, tag, i)
            _gn_bottleneck(s, p, ci, co)
            if ci != co:
                _conv(s, p + ".downsample.0", co, ci, 1); _gn(s, p + ".downsample.1", co)
    _gct(s, h + ".GCT_sc", 256 + DEC); _conv(s, h + ".conv_sc", REFINE, 256 + DEC, 1); _gn(s, h + ".bn_sc", REFINE)
    _lin(s, h + ".IA10.IA", DEC + REFINE, HEAD + DEC + REFINE)
    _conv(s, h + ".conv1", DEC // 2, DEC + REFINE, 3); _gn(s, h + ".bn1", DEC // 2)
    _lin(s, h + ".IA11.IA", DEC // 2, HEAD + DEC // 2)
    _conv(s, h + ".conv2", DEC // 2, DEC // 2, 3); _gn(s, h + ".bn2", DEC // 2)
    _lin(s, h + ".IA_final_fg", DEC // 2 + 1, HEAD)
    _lin(s, h + ".IA_final_bg", DEC // 2 + 1, HEAD)
    _conv(s, "dynamic_prehead.conv", PRE, PREHEAD_IN, 1, bias=True); _gn(s, "dynamic_prehead.bn", PRE)
    return s


_BUFFER_KINDS = ("bn_weight", "bn_bias", "bn_running_mean", "bn_running_var")  # FrozenBatchNorm2d buffers
_ALIASES = {"semantic_embedding.0": "seperate_conv", "semantic_embedding.1": "bn1",
            "semantic_embedding.3": "embedding_conv", "semantic_embedding.4": "bn2"}


class ParamTree(nn.Module):
    """nn.Module whose state_dict() has exactly the reference's names (values default-initialised)."""

    def __init__(self):
        super().__init__()
        spec = param_spec()
        for name, (kind, shape) in spec.items():
            if name.startswith("semantic_embedding."):
                continue
            t = _default_value(kind, shape)
            mod, leaf = self._descend(name)
            if kind in _BUFFER_KINDS:
                mod.register_buffer(leaf, t)
            else:
                mod.register_parameter(leaf, nn.Parameter(t, requires_grad=False))
        # the reference registers the embedding head twice (attributes + an nn.Sequential of the
        # same modules, aocnet.py:19-25); the aliases share storage.
        seq = nn.Module()
        for alias, target in _ALIASES.items():
            seq.add_module(alias.split(".")[1], getattr(self, target))
        self.add_module("semantic_embedding", seq)

    def _descend(self, dotted):
        parts = dotted.split(".")
        mod = self
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        return mod, parts[-1]


def _default_value(kind, shape):
    if kind in ("bn_weight", "bn_running_var", "gn_weight", "gct_alpha"):
        return torch.ones(shape)
    return torch.zeros(shape)


def synthetic_state_dict(seed=1234):
    """Seeded random weights with every path active (non-trivial GCT gates, biases, BN statistics).
    Scales are chosen so activations stay O(1..10) through the 101-layer backbone without a
    calibration pass (stock initialisers reach 3.6e5)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, (kind, shape) in param_spec().items():
        if name.startswith("semantic_embedding."):
            continue
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif kind == "linear":
            t = torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind in ("gn_weight", "bn_weight"):
            t = 1.0 + 0.2 * torch.randn(shape, generator=g)
            if kind == "bn_weight" and name.endswith("bn3.weight"):
                t = 0.3 * t  # keeps the 33 residual adds of ResNet101 from blowing activations up
        elif kind in ("gn_bias", "bn_bias"):
            t = 0.2 * torch.randn(shape, generator=g)
        elif kind == "gct_alpha":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind in ("gct_gamma", "gct_beta"):
            t = 0.3 * torch.randn(shape, generator=g)
        elif kind == "dis_bias":
            t = torch.randn(shape, generator=g) * 0.5
        elif kind == "bn_running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_running_var":
            t = 0.8 + 0.4 * torch.rand(shape, generator=g)
        else:
            raise KeyError(kind)
        sd[name] = t
    for alias, target in _ALIASES.items():
        for leaf in ("weight", "bias"):
            sd[alias + "." + leaf] = sd[target + "." + leaf]
    # order like the reference
    return OrderedDict((k, sd[k]) for k in param_spec().keys())


def frozen_bn_names():
    return [n[:-len(".running_mean")] for n, (k, _) in param_spec().items() if k == "bn_running_mean"]
