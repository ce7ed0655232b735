"""tcgen05 (3xTF32) kernels against fp64 / the fp32 SIMT kernels of the same library."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import from_T, report, to_T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(state_dict):
    from aocb200.engine import Engine
    return Engine(state_dict, torch.device("cuda:0"))


def test_gemm_tf32x3(eng):
    g = torch.Generator().manual_seed(0)
    M, N, K = 256, 768, 100
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    want = (A.double() @ B.double().t())
    ws = torch.empty(eng.L.tc_image_bytes(M, 104, 128) + eng.L.tc_image_bytes(N, 104, 256), dtype=torch.uint8, device="cuda")
    res = {}
    for variant in (0, 1):
        C = torch.full((M, N), float("nan"), device="cuda")
        eng.L.gemm_tf32x3_test(A.cuda().data_ptr(), B.cuda().data_ptr(), C.data_ptr(), M, N, K, variant, ws.data_ptr(),
                               ws.numel(), eng.stream)
        torch.cuda.synchronize()
        err = (C.cpu().double() - want).abs().max().item()
        res[variant] = err
        print("[parity] tcgen05 3xTF32 gemm variant %d: max|d|=%.3e (|C| max %.2f; fp32 matmul err %.3e)" %
              (variant, err, want.abs().max().item(), (A @ B.t()).double().sub(want).abs().max().item()))
    assert res[0] < 2e-5, res


def test_conv_tc_vs_torch(eng):
    from test_gpu_ops import _conv_case
    eng.tc_conv = True
    try:
        _conv_case(eng, 1, 33, 41, 4, 64, 7, 2, 3, 1, True, False, False, True, seed=101, tolmul=3.0)
        _conv_case(eng, 1, 17, 23, 64, 256, 1, 1, 0, 1, True, True, False, True, seed=102, tolmul=3.0)
        _conv_case(eng, 1, 17, 23, 128, 128, 3, 2, 1, 1, True, False, False, True, seed=103, tolmul=3.0)
        _conv_case(eng, 1, 9, 13, 512, 512, 3, 1, 4, 4, True, False, False, True, seed=104, tolmul=3.0)
        _conv_case(eng, 3, 19, 21, 164, 64, 1, 1, 0, 1, False, False, True, False, seed=105, tolmul=3.0)
        _conv_case(eng, 2, 19, 21, 256, 100, 1, 1, 0, 1, False, False, False, True, seed=106, tolmul=3.0)
        _conv_case(eng, 2, 13, 17, 24, 64, 1, 1, 0, 1, False, False, False, True, seed=107, tolmul=3.0)
        _conv_case(eng, 6, 61, 107, 320, 128, 3, 1, 1, 1, False, False, False, False, seed=108, tolmul=3.0)
        _conv_case(eng, 2, 1, 1, 512, 128, 1, 1, 0, 1, True, False, False, False, seed=109, tolmul=3.0)
        _conv_case(eng, 2, 12, 14, 48, 64, 3, 1, 6, 6, False, False, True, False, ld_in=80, off_in=16, seed=110, tolmul=3.0)
        _conv_case(eng, 1, 61, 107, 2048, 256, 3, 1, 6, 6, True, False, False, True, seed=111, tolmul=3.0)
        _conv_case(eng, 6, 61, 107, 256, 512, 1, 1, 0, 1, False, False, False, False, seed=112, tolmul=3.0)
    finally:
        eng.tc_conv = True


def test_global_match_tc_vs_simt(eng):
    from test_gpu_ops import _rand_scene
    for seed, h, w, K, F_, absent in ((1, 25, 33, 3, 2, None), (2, 61, 107, 5, 2, 4), (3, 33, 37, 1, 3, None)):
        embs, masks, _ = _rand_scene(seed, h, w, K, F_, absent)
        outs = {}
        for tc in (False, True):
            eng.tc_match = tc
            eng.bank.reset()
            eng.keep_debug = True
            np.random.seed(seed)
            eng.match_features([e.cuda() for e in embs[:F_]], [m.cuda() for m in masks[:F_]], embs[F_].cuda(),
                               masks[F_].cuda(), to_T(embs[F_ + 1], eng), K)
            torch.cuda.synchronize()
            outs[tc] = eng.debug["g"].clone().cpu()
        eng.keep_debug = False
        eng.tc_match = True
        report("global match tcgen05 vs simt (seed %d)" % seed, outs[True], outs[False], 5e-6)
