import ctypes, sys, os, torch
sys.path.insert(0, "/root/repo")
def run(libpath, new_api, shapes, opt=None):
    L = ctypes.CDLL(libpath)
    if opt: print(opt, L.aoc_set_option(opt[0], opt[1]))
    L.aoc_conv_packed_weight_bytes.restype = ctypes.c_size_t
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    for name, N, H, W, Cin, Cout, k, pad in shapes:
        x = torch.randn(N * H * W * Cin, generator=g).to(dev)
        w = (torch.randn(Cout, k, k, Cin, generator=g) / (Cin * k * k) ** 0.5).to(dev)
        nb = L.aoc_conv_packed_weight_bytes(Cout, Cin, k, k)
        wp = torch.empty(nb, dtype=torch.uint8, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        L.aoc_conv_pack_weights_tf32x3(P(w), Cout, Cin, k, k, P(wp), st)
        y = torch.empty(N * H * W * Cout, device=dev)
        wsb = torch.empty(8 * N * H * W * Cout * 4 if N * H * W <= 128 * 74 else 16, dtype=torch.uint8, device=dev)
        def call():
            args = [P(x), P(wp), None, None, None, None, 0, P(y), None, N, H, W, Cin, Cin, Cout, Cout, 0, k, k, 1, pad, 1, 0, 0]
            if new_api: args += [P(wsb), ctypes.c_size_t(wsb.numel())]
            args.append(st)
            rc = L.aoc_conv2d_nhwc_tc(*args)
            assert rc == 0, rc
        for _ in range(3): call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): call()
        e1.record(); torch.cuda.synchronize()
        print("%-28s %-40s %8.1f us" % (os.path.basename(libpath), name, 100 * e0.elapsed_time(e1)))
shapes = [("bb.l3 256->1024 1x1", 1, 31, 54, 256, 1024, 1, 0), ("dec 256->512 1x1 x6", 6, 61, 107, 256, 512, 1, 0),
          ("dec 164->256 1x1 x6", 6, 121, 213, 164, 256, 1, 0), ("emb 256->100 1x1", 1, 121, 213, 256, 100, 1, 0),
          ("dec.l1.conv3 64->256 1x1 x6", 6, 121, 213, 64, 256, 1, 0), ("dec.l1.conv1 164->64 1x1 x6", 6, 121, 213, 164, 64, 1, 0),
          ("dec.half 128->512 1x1 x6", 6, 61, 107, 128, 512, 1, 0), ("bb.l1 64->256 1x1", 1, 121, 213, 64, 256, 1, 0),
          ("dec.conv2 128->128 3x3 x6", 6, 121, 213, 128, 128, 3, 1),
          ("bb.l3 1024->256 1x1", 1, 31, 54, 1024, 256, 1, 0), ("bb.l3 256->256 3x3", 1, 31, 54, 256, 256, 3, 1),
          ("bb.l3 256->1024 1x1", 1, 31, 54, 256, 1024, 1, 0), ("bb.l4 512->512 3x3", 1, 31, 54, 512, 512, 3, 1),
          ("bb.l4 2048->512 1x1", 1, 31, 54, 2048, 512, 1, 0), ("bb.l4 512->2048 1x1", 1, 31, 54, 512, 2048, 1, 0),
          ("bb.aspp 2048->256 3x3", 1, 31, 54, 2048, 256, 3, 1), ("bb.aspp.conv1 1280->256 1x1", 1, 31, 54, 1280, 256, 1, 0),
          ("bb.l2 512->128 1x1", 1, 61, 107, 512, 128, 1, 0), ("bb.l2 128->128 3x3", 1, 61, 107, 128, 128, 3, 1),
          ("bb.l2 128->512 1x1", 1, 61, 107, 128, 512, 1, 0),
          ("dec.conv1 320->128 3x3 x6", 6, 121, 213, 320, 128, 3, 1), ("dec.l1.conv2 64->64 3x3 x6", 6, 121, 213, 64, 64, 3, 1),
          ("dec.aspp-like 512->128 3x3 half x6", 6, 61, 107, 512, 128, 3, 1)]
run("/root/repo/aocb200/libaocb200_prev.so", True, shapes)
run("/root/repo/aocb200/libaocb200.so", True, shapes)
run("/root/repo/aocb200/libaocb200.so", True, shapes, (b"conv_narrow_nit", 0))
