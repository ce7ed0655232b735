"""Synthetic clips in the reference's input contract (no datasets are available offline).

`make_clip` follows SURVEY.md section 8d: smooth background, K moving textured ellipses, first-frame
uint8 label map, frames normalised like MultiToTensor (dataloaders/custom_transforms.py:479-481:
/255, ImageNet mean/std, HWC->CHW).
"""
import numpy as np
import torch

_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def _smooth_noise(rng, H, W, cells=12):
    gh, gw = max(2, H // cells), max(2, W // cells)
    g = rng.uniform(0, 255, size=(gh, gw, 3)).astype(np.float32)
    t = torch.from_numpy(g).permute(2, 0, 1)[None]
    t = torch.nn.functional.interpolate(t, size=(H, W), mode="bicubic", align_corners=True)
    return t[0].permute(1, 2, 0).numpy().clip(0, 255)


def make_clip(seed, H, W, K, T):
    """Returns (frames [T,3,H,W] float32 normalised, labels [T,H,W] uint8 with ids 0..K)."""
    rng = np.random.default_rng(seed)
    bg = _smooth_noise(rng, H, W)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    m = min(H, W)
    objs = []
    for k in range(K):
        objs.append(dict(
            cy=rng.uniform(0.2, 0.8) * H, cx=rng.uniform(0.15, 0.85) * W,
            ay=rng.uniform(0.08, 0.2) * m, ax=rng.uniform(0.08, 0.2) * m,
            vy=rng.uniform(-3, 3), vx=rng.uniform(-3, 3),
            col=rng.uniform(30, 225, size=3).astype(np.float32),
            tex=rng.normal(0, 12, size=(H, W, 3)).astype(np.float32)))
    frames = np.empty((T, 3, H, W), dtype=np.float32)
    labels = np.zeros((T, H, W), dtype=np.uint8)
    for t in range(T):
        img = bg.copy()
        for k, o in enumerate(objs):
            cy, cx = o["cy"] + t * o["vy"], o["cx"] + t * o["vx"]
            inside = ((yy - cy) / o["ay"]) ** 2 + ((xx - cx) / o["ax"]) ** 2 <= 1.0
            img[inside] = (o["col"] + o["tex"])[inside]
            labels[t][inside] = k + 1
        img = img + rng.normal(0, 2, size=img.shape).astype(np.float32)
        img = (img.clip(0, 255) / 255.0 - _MEAN) / _STD
        frames[t] = img.transpose(2, 0, 1)
    return torch.from_numpy(frames), torch.from_numpy(labels)


def restrict_size(H, W, max_size=1040):
    """MultiRestrictSize (dataloaders/custom_transforms.py:395-430): long edge <= max_size, then
    H-1 and W-1 rounded to multiples of 16."""
    if max(H, W) > max_size:
        sc = float(max_size) / max(H, W)
        H, W = H * sc, W * sc
    H = int(np.around((H - 1) / 16.0) * 16 + 1)
    W = int(np.around((W - 1) / 16.0) * 16 + 1)
    return H, W
