"""Writes tests/golden/io_edges.npz from the REFERENCE's own transforms (run in the build container, where /root/reference
is mounted): dataloaders/custom_transforms.py MultiRestrictSize + MultiToTensor on random uint8 frames (sizes, mirrored
twins, resized + normalised tensors) and the `_palette` list of utils/image.py.  The fixtures are what
tests/test_io_cpu.py and tests/test_gpu_ops.py::test_prepare_frame compare aocb200/io.py with."""
import importlib.util
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/AOC-Net/complete_project/AOCNet"


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ct = load(os.path.join(REF, "dataloaders", "custom_transforms.py"), "ref_custom_transforms")
    out = {}
    # --- size arithmetic over many shapes / settings
    rs = np.random.RandomState(0)
    cases = []
    for (h, w) in [(480, 854), (720, 1280), (1080, 1920), (360, 640), (854, 480), (481, 849), (100, 37), (17, 17), (600, 600)] + \
                  [tuple(int(v) for v in rs.randint(20, 1400, 2)) for _ in range(30)]:
        for kw in (dict(max_size=800 * 1.3, multi_scale=[1.0]), dict(max_size=800, multi_scale=[1.3]),
                   dict(max_size=1040, multi_scale=[1.0, 1.15, 1.3], flip=True), dict(min_size=480, max_size=None, multi_scale=[1.0, 0.75], flip=True)):
            tr = ct.MultiRestrictSize(**kw)
            img = np.zeros((h, w, 3), np.float32)
            smp = tr({"current_img": img, "meta": {"flip": False}})
            cases.append([h, w, -1 if kw.get("min_size") is None else kw["min_size"], -1 if kw.get("max_size") is None else kw["max_size"],
                          len(kw["multi_scale"]), int(bool(kw.get("flip")))] + list(kw["multi_scale"]) + [0] * (3 - len(kw["multi_scale"])) +
                         [v for s in smp for v in (s["current_img"].shape[0], s["current_img"].shape[1], int(s["meta"].get("flip", False)))])
    width = max(len(c) for c in cases)
    out["size_cases"] = np.array([c + [-1] * (width - len(c)) for c in cases], dtype=np.float64)
    # --- resize + mirror + MultiToTensor on small random frames
    k = 0
    for (h, w, kw) in [(70, 90, dict(max_size=64, multi_scale=[1.0, 1.3], flip=True)), (95, 61, dict(max_size=80, multi_scale=[1.0])),
                       (49, 65, dict(max_size=800, multi_scale=[1.0], flip=True))]:
        img_u8 = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        sample = {"current_img": np.array(img_u8, dtype=np.float32), "meta": {"flip": False}}     # dataloaders/datasets.py: float32 frames
        smp = ct.MultiToTensor()(ct.MultiRestrictSize(**kw)(sample))
        out["frame%d_u8" % k] = img_u8
        out["frame%d_kw" % k] = np.array([kw["max_size"], len(kw["multi_scale"]), int(bool(kw.get("flip")))] + list(kw["multi_scale"]), dtype=np.float64)
        for a, s in enumerate(smp):
            out["frame%d_aug%d" % (k, a)] = s["current_img"].numpy().astype(np.float32)
            out["frame%d_aug%d_flip" % (k, a)] = np.array(int(s["meta"].get("flip", False)))
        out["frame%d_n" % k] = np.array(len(smp))
        k += 1
    out["n_frames"] = np.array(k)
    # --- palette
    src = open(os.path.join(REF, "utils", "image.py")).read()
    m = re.search(r"^_palette\s*=\s*\[([^\]]*)\]", src, flags=re.M)
    pal = np.array([int(v) for v in m.group(1).split(",")], dtype=np.uint8)
    assert pal.size == 768
    out["palette"] = pal.reshape(256, 3)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "io_edges.npz"), **out)
    print("wrote io_edges.npz:", len(cases), "size cases,", k, "frames")


if __name__ == "__main__":
    main()
