"""The image edges either side of the eval loop: what the reference's dataset transforms do to a frame
before `forward_for_eval` sees it, and what its evaluator does with a predicted label map afterwards.

    multi_restrict_size   dataloaders/custom_transforms.py:387-463 (MultiRestrictSize): the size arithmetic -- long edge
                          <= max_size (or short edge <= min_size), every scale of `multi_scale`, H-1 and W-1 rounded to
                          multiples of 16, the mirrored twin of each scale with flip
    prepare_frame         the same transform's cv2.resize(INTER_CUBIC) + mirror, then MultiToTensor (:465-487: /255, -mean,
                          /std, HWC -> CHW) as ONE CUDA kernel on the uint8 frame (aoc_prepare_frame_u8, csrc/resize.cu)
    prepare_samples       frame (+ label) -> the list of augmentations `aocb200.sequence.run_sequence_tta` consumes
    encode_mask_png /     utils/image.py:40-44 (save_mask): uint8 label map -> 8-bit palette PNG with the DAVIS palette
    save_mask_png         (written here with zlib: no PIL on the path; PIL reads it back identically, tests/test_io_cpu.py)

JPEG decoding stays with the caller (the reference reads frames with cv2.imread on the host, dataloaders/datasets.py) and
the robustness benchmark's perturbations (Robust-VOS-Benchmark/.../datasets_robustness.py:459-506) are dataset generation,
not evaluation: both are outside the per-frame path and are not rebuilt.
"""
import struct
import zlib

import numpy as np
import torch

MEAN = (0.485, 0.456, 0.406)          # custom_transforms.py:478-481
STD = (0.229, 0.224, 0.225)


def multi_restrict_size(h, w, min_size=None, max_size=800, multi_scale=(1.3,), flip=False):
    """-> [dict(h, w, flip)] in the order MultiRestrictSize emits its samples (per scale: plain, then mirrored)."""
    assert (min_size is None) or (max_size is None)
    out = []
    for scale in multi_scale:
        sc = None
        if min_size is not None:
            short_edge = w if h > w else h
            if short_edge > min_size:
                sc = float(min_size) / short_edge
        else:
            long_edge = h if h > w else w
            if long_edge > max_size:
                sc = float(max_size) / long_edge
        new_h, new_w = (h, w) if sc is None else (sc * h, sc * w)
        new_h, new_w = int(new_h * scale), int(new_w * scale)
        if (new_h - 1) % 16 != 0:
            new_h = int(np.around((new_h - 1) / 16.) * 16 + 1)
        if (new_w - 1) % 16 != 0:
            new_w = int(np.around((new_w - 1) / 16.) * 16 + 1)
        out.append(dict(h=new_h, w=new_w, flip=False))
        if flip:
            out.append(dict(h=new_h, w=new_w, flip=True))
    return out


def prepare_frame(img_u8, size=None, flip=False, device=None, out=None):
    """img_u8: uint8 [H, W, 3] (numpy or torch, host or device; channel order as the caller read it) -> float32
    [1, 3, h, w] on the device: resized to `size` = (h, w) like cv2.resize(INTER_CUBIC) on the float image when it differs
    from (H, W), mirrored when `flip`, normalised like MultiToTensor.  One kernel; the uint8 frame is the only H2D copy."""
    from aocb200.lib import lib
    t = torch.as_tensor(img_u8)
    assert t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 3, "uint8 [H, W, 3] expected"
    dev = torch.device(device) if device is not None else (t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    t = t.contiguous().to(dev, non_blocking=True)
    H, W = int(t.shape[0]), int(t.shape[1])
    h, w = (H, W) if size is None else (int(size[0]), int(size[1]))
    if out is None:
        out = torch.empty((1, 3, h, w), dtype=torch.float32, device=dev)
    assert out.is_contiguous() and tuple(out.shape) == (1, 3, h, w)
    import ctypes
    mean, std = (ctypes.c_float * 3)(*MEAN), (ctypes.c_float * 3)(*STD)
    with torch.cuda.device(dev):
        lib().prepare_frame_u8(t.data_ptr(), H, W, h, w, 1 if flip else 0, mean, std, out.data_ptr(),
                               torch.cuda.current_stream(dev).cuda_stream)
    return out


def prepare_samples(img_u8, label=None, min_size=None, max_size=800, multi_scale=(1.3,), flip=False, device=None):
    """One frame -> the augmentation list of the reference loop (eval_manager_mm.py:212-245): dict(img [1,3,h,w] device
    float, label [H,W] or None -- labels are never resized by the transform, only mirrored --, flip)."""
    t = torch.as_tensor(img_u8)
    H, W = int(t.shape[0]), int(t.shape[1])
    samples = []
    for s in multi_restrict_size(H, W, min_size, max_size, multi_scale, flip):
        lab = None
        if label is not None:
            lab = torch.as_tensor(label)
            if s["flip"]:
                lab = torch.flip(lab, dims=[lab.dim() - 1])
        samples.append(dict(img=prepare_frame(t, (s["h"], s["w"]), s["flip"], device), label=lab, flip=s["flip"]))
    return samples


def davis_palette():
    """The 256-entry palette of utils/image.py:14 (`_palette`): the PASCAL-VOC colour map for ids 0..21 with the DAVIS
    convention 191 in place of 192, grey (i, i, i) from id 22 on.  -> uint8 [256, 3]"""
    pal = np.zeros((256, 3), dtype=np.uint8)
    for i in range(22):
        c, r, g, b = i, 0, 0, 0
        for j in range(8):
            r |= ((c >> 0) & 1) << (7 - j)
            g |= ((c >> 1) & 1) << (7 - j)
            b |= ((c >> 2) & 1) << (7 - j)
            c >>= 3
        pal[i] = [191 if v == 192 else v for v in (r, g, b)]
    for i in range(22, 256):
        pal[i] = (i, i, i)
    return pal


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)


def encode_mask_png(label, palette=None, level=6):
    """uint8 [H, W] label map -> bytes of an 8-bit indexed-colour PNG (colour type 3, PLTE = palette), the file format
    save_mask produces through PIL (`Image.fromarray(mask).convert('P')`, `putpalette`, `save`)."""
    m = np.ascontiguousarray(np.asarray(label, dtype=np.uint8))
    assert m.ndim == 2
    H, W = m.shape
    pal = davis_palette() if palette is None else np.asarray(palette, dtype=np.uint8).reshape(-1, 3)
    raw = np.empty((H, W + 1), dtype=np.uint8)
    raw[:, 0] = 0                                   # filter type 0 (None) on every scan line
    raw[:, 1:] = m
    ihdr = struct.pack(">IIBBBBB", W, H, 8, 3, 0, 0, 0)
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + _chunk(b"PLTE", pal.tobytes()) +
            _chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + _chunk(b"IEND", b""))


def save_mask_png(label, path, palette=None):
    """label: uint8 [H, W] tensor (device or host) or array -- e.g. DeviceSequence.step()'s result -- written to `path`."""
    if torch.is_tensor(label):
        label = label.detach().to("cpu", torch.uint8).numpy()
    data = encode_mask_png(label, palette)
    with open(path, "wb") as f:
        f.write(data)
    return len(data)
