"""Robustness of the mbarrier / named-barrier protocols: the conv kernel's two-issuer hand-over, the
fused linear-attention kernel's five-role pipeline and the tcgen05 attention kernel are launched hundreds of times back to
back, with no host synchronisation in between, over shapes that exercise both issuer modes, ragged tiles, multi-image tiles
and persistent CTAs with many tiles.  A protocol slip shows up as a trapped launch (bounded waits, ptx.cuh) or as bits that
differ from the first launch: every kernel here is required to be bitwise repeatable."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def test_unet_forward_300_back_to_back():
    """300 U-Net evaluations alternating between three engines (32x32 B=8: issuer mode 2 + folded kernels + fused linear
    attention; 16x24 B=3: ragged tiles; 64x64 B=2: many tiles per CTA), each compared bitwise with its first result."""
    import diffusion_models_b200 as ddm
    from oracle import synth_state_dict
    model = ddm.Unet(dim=64, dim_mults=(1, 2, 4, 8))
    model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 0))
    model = model.cuda().eval()
    cases = [(8, 32, 32), (3, 16, 24), (2, 64, 64)]
    xs = [rnd((b, 3, h, w), 400 + i).cuda() for i, (b, h, w) in enumerate(cases)]
    ts = [torch.full((b,), 500 - 100 * i, device="cuda") for i, (b, _, _) in enumerate(cases)]
    first = [model(x, t) for x, t in zip(xs, ts)]
    assert all(torch.isfinite(f).all() for f in first)
    n0 = ddm._lib.launch_count()
    outs = []
    for it in range(300):
        i = it % 3
        outs.append((i, model(xs[i], ts[i])))          # eager launches, no synchronisation between forwards
        if len(outs) == 30:
            for j, y in outs:
                assert torch.equal(y, first[j]), f"forward {it} of engine {j} differs from its first run"
            outs.clear()
    assert ddm._lib.launch_count() - n0 > 300 * 80


def test_conv_issuer_modes_back_to_back():
    """Both two-issuer modes of conv_tc_kernel (alternate tiles with split rings; alternate stages with the ordering token),
    streamed and resident weights, ragged tiles: 200 launches each without a host sync, all bitwise equal."""
    from diffusion_models_b200 import _lib
    from diffusion_models_b200._lib import ConvArgs
    from diffusion_models_b200.packing import pack_conv
    lib = _lib.init(0)
    s = torch.cuda.current_stream().cuda_stream
    for (B, H, W, Cin, Cout, seed) in [(64, 32, 32, 64, 64, 1),      # resident weights, mode 2 (a tile fits half the ring), 512 tiles
                                       (32, 32, 32, 128, 64, 2),     # streamed weights, mode 1 (stage alternation)
                                       (16, 16, 16, 192, 128, 3),    # streamed, 27 k-chunks per tile
                                       (7, 10, 20, 64, 64, 4),       # ragged in y and x
                                       (33, 4, 4, 512, 512, 5)]:     # two N tiles, batch-packed tiles, ragged batch
        x = rnd((B, H, W, Cin), 500 + seed).to("cuda", BF)
        pk = pack_conv(rnd((Cout, Cin, 3, 3), 510 + seed, (Cin * 9) ** -0.5))
        w = pk.weight.cuda()
        bias = rnd((Cout,), 520 + seed, 0.1).cuda()
        outs = [torch.zeros((B, H, W, Cout), dtype=BF, device="cuda") for _ in range(2)]
        a = ConvArgs()
        a.src0, a.C0, a.ld0, a.view = x.data_ptr(), Cin, Cin, 0
        a.B, a.H, a.W, a.ntaps = B, H, W, 9
        for i, (dy, dx, p) in enumerate(pk.taps):
            a.tap_dy[i], a.tap_dx[i], a.tap_p[i] = dy, dx, p
        a.weight, a.N, a.N_pad, a.K_pad = w.data_ptr(), pk.n, pk.n_pad, pk.k_pad
        a.bias, a.act = bias.data_ptr(), 1
        a.ld_out, a.OH, a.OW, a.sy, a.sx = Cout, H, W, 1, 1
        a.out = outs[0].data_ptr()
        _lib.check(lib.ddm_conv2d(C.byref(a), s))
        a.out = outs[1].data_ptr()
        for _ in range(200):
            _lib.check(lib.ddm_conv2d(C.byref(a), s))
        torch.cuda.synchronize()
        assert torch.isfinite(outs[0].float()).all() and torch.equal(outs[0], outs[1]), (B, H, W, Cin, Cout)


def test_cta_pair_norm_and_split_k_back_to_back():
    """The cluster protocol of the 512-channel Block tail (two CTAs exchanging row sums of squares through distributed shared
    memory, double-buffered by tile parity) and the split-K instantiation: 200 launches each without a host sync, bitwise equal.
    2400 images of 4 x 4 = 300 M tiles: every cluster walks four or five tiles; 5 images: a single, ragged tile pair."""
    from diffusion_models_b200 import _lib
    from diffusion_models_b200._lib import ConvArgs
    from diffusion_models_b200.packing import pack_conv
    lib = _lib.init(0)
    s = torch.cuda.current_stream().cuda_stream
    for (B, Cin, seed) in [(2400, 128, 1), (5, 512, 2)]:
        H = W = 4
        Cout = 512
        x = rnd((B, H, W, Cin), 600 + seed).to("cuda", BF)
        pk = pack_conv(rnd((Cout, Cin, 3, 3), 610 + seed, (Cin * 9) ** -0.5))
        w = pk.weight.cuda()
        bias, g = rnd((Cout,), 620 + seed, 0.1).cuda(), (1 + 0.1 * rnd((Cout,), 630 + seed)).cuda() * Cout ** 0.5
        ss, res = rnd((1, 2 * Cout), 640 + seed, 0.3).cuda(), rnd((B, H, W, Cout), 650 + seed).to("cuda", BF)
        outs = [torch.zeros((B, H, W, Cout), dtype=BF, device="cuda") for _ in range(2)]
        a = ConvArgs()
        a.src0, a.C0, a.ld0 = x.data_ptr(), Cin, Cin
        a.B, a.H, a.W, a.ntaps = B, H, W, 9
        for i, (dy, dx, p) in enumerate(pk.taps):
            a.tap_dy[i], a.tap_dx[i], a.tap_p[i] = dy, dx, p
        a.weight, a.N, a.N_pad, a.K_pad = w.data_ptr(), pk.n, pk.n_pad, pk.k_pad
        a.bias, a.norm_g, a.scale_shift, a.act = bias.data_ptr(), g.data_ptr(), ss.data_ptr(), 1
        a.residual, a.ld_res = res.data_ptr(), Cout
        a.ld_out, a.OH, a.OW, a.sy, a.sx = Cout, H, W, 1, 1
        a.out = outs[0].data_ptr()
        _lib.check(lib.ddm_conv2d(C.byref(a), s))
        a.out = outs[1].data_ptr()
        for _ in range(200):
            _lib.check(lib.ddm_conv2d(C.byref(a), s))
        torch.cuda.synchronize()
        assert torch.isfinite(outs[0].float()).all() and torch.equal(outs[0], outs[1]), (B, Cin)
    # split-K: 16 images of 4 x 4 (two M tiles), 512 -> 512, seven K ranges
    B, H, W, Cin, Cout, ks = 16, 4, 4, 512, 512, 7
    x = rnd((B, H, W, Cin), 660).to("cuda", BF)
    pk = pack_conv(rnd((Cout, Cin, 3, 3), 661, (Cin * 9) ** -0.5))
    w = pk.weight.cuda()
    parts = [torch.zeros((ks, B * H * W, Cout), dtype=torch.float32, device="cuda") for _ in range(2)]
    a = ConvArgs()
    a.src0, a.C0, a.ld0 = x.data_ptr(), Cin, Cin
    a.B, a.H, a.W, a.ntaps = B, H, W, 9
    for i, (dy, dx, p) in enumerate(pk.taps):
        a.tap_dy[i], a.tap_dx[i], a.tap_p[i] = dy, dx, p
    a.weight, a.N, a.N_pad, a.K_pad = w.data_ptr(), pk.n, pk.n_pad, pk.k_pad
    a.ld_out, a.OH, a.OW, a.sy, a.sx = Cout, H, W, 1, 1
    a.ksplit, a.partial_out = ks, parts[0].data_ptr()
    _lib.check(lib.ddm_conv2d(C.byref(a), s))
    a.partial_out = parts[1].data_ptr()
    for _ in range(200):
        _lib.check(lib.ddm_conv2d(C.byref(a), s))
    torch.cuda.synchronize()
    assert torch.isfinite(parts[0]).all() and torch.equal(parts[0], parts[1])


def test_fused_linear_attention_and_attention_back_to_back():
    from diffusion_models_b200 import _lib
    from diffusion_models_b200._lib import LinAttnBlockArgs
    from diffusion_models_b200.packing import linattn_k_shift, norm_gain, pack_conv
    lib = _lib.init(0)
    s = torch.cuda.current_stream().cuda_stream
    heads, d, hid = 4, 32, 128
    for B, n in ((600, 128), (300, 1024), (5, 4096)):      # 1 tile per pass (T = 1), several images per CTA, many tiles per image
        x = rnd((B, n, 64), 600).to("cuda", BF)
        w_qkv, g_in, w_out, mem = rnd((3 * hid, 64, 1, 1), 601, 0.125), torch.ones((1, 64, 1, 1)), rnd((64, hid, 1, 1), 602, 0.09), rnd((2, heads, d, 4), 603)
        keep = [pack_conv(w_qkv, in_scale=norm_gain(g_in)).weight.cuda(), pack_conv(w_out).weight.cuda(), torch.zeros(64, device="cuda"),
                norm_gain(torch.ones(1, 64, 1, 1)).cuda(), mem.cuda(), linattn_k_shift(w_qkv, g_in, mem, heads, d).cuda()]
        outs = [torch.zeros_like(x) for _ in range(2)]
        a = LinAttnBlockArgs()
        a.x, a.B, a.n, a.C = x.data_ptr(), B, n, 64
        a.w_qkv, a.w_out, a.bias_out, a.g_out, a.mem_kv, a.k_shift = (t.data_ptr() for t in keep)
        a.heads, a.dim_head, a.n_mem = heads, d, 4
        a.out = outs[0].data_ptr()
        _lib.check(lib.ddm_linear_attention_block(C.byref(a), s))
        a.out = outs[1].data_ptr()
        for _ in range(100):
            _lib.check(lib.ddm_linear_attention_block(C.byref(a), s))
        torch.cuda.synchronize()
        assert torch.isfinite(outs[0].float()).all() and torch.equal(outs[0], outs[1]), (B, n)
    for B, nq, nk, hh, dd, n_mem in ((700, 16, 16, 4, 32, 4), (40, 64, 77, 4, 32, 0), (9, 256, 256, 1, 128, 0)):
        q, k, v = (rnd((B, n_, hh * dd), 610 + i).to("cuda", BF) for i, n_ in enumerate((nq, nk, nk)))
        mk, mv = (rnd((hh, max(n_mem, 1), dd), 615 + i).cuda() for i in range(2))
        outs = [torch.zeros((B, nq, hh * dd), dtype=BF, device="cuda") for _ in range(2)]
        for i, reps in ((0, 1), (1, 100)):
            for _ in range(reps):
                _lib.check(lib.ddm_attention(q.data_ptr(), hh * dd, k.data_ptr(), hh * dd, v.data_ptr(), hh * dd, mk.data_ptr() if n_mem else None,
                                             mv.data_ptr() if n_mem else None, n_mem, outs[i].data_ptr(), B, nq, nk, hh, dd, s))
        torch.cuda.synchronize()
        assert torch.isfinite(outs[0].float()).all() and torch.equal(outs[0], outs[1]), (B, nq, nk)
