"""Per-kernel parity on a B200: every C-ABI entry point of include/ddm_b200.h against tests/kernel_ref.py on random
inputs (bit-level rounding aside: bf16 outputs are compared with rtol/atol of one bf16 ulp-ish, 1e-2)."""
import ctypes as C

import pytest
import torch

import kernel_ref as R

pytestmark = pytest.mark.gpu

BF, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def lib():
    from diffusion_models_b200 import _lib
    return _lib.init(0)


def dev(t, dtype=None):
    return t.to("cuda", dtype or t.dtype).contiguous()


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def stream():
    return torch.cuda.current_stream().cuda_stream


def check(code):
    from diffusion_models_b200 import _lib
    _lib.check(code)
    torch.cuda.synchronize()


def close(a, b, tol=1e-2):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs()
    bound = tol * b.abs() + tol
    bad = (err > bound).float().mean().item()
    assert bad < 1e-4, f"{bad * 100:.3f}% elements off; max err {err.max().item():.4f}; rel-l2 {(a - b).norm() / b.norm():.4e}"


def run_conv(lib, srcs, pk, domain, out, **kw):
    """srcs: list of device bf16 tensors [B,Hs,Ws,C]; pk: PackedConv.  Returns (kernel out, reference out)."""
    from diffusion_models_b200._lib import ConvArgs
    w = dev(pk.weight)
    a = ConvArgs()
    a.src0 = srcs[0].data_ptr()
    a.src1 = srcs[1].data_ptr() if len(srcs) > 1 else None
    a.C0 = pk.seg_channels[0] // (2 if pk.view else 1)
    a.C1 = pk.seg_channels[1] if len(srcs) > 1 else 0
    a.ld0 = srcs[0].shape[-1]
    a.ld1 = srcs[1].shape[-1] if len(srcs) > 1 else 0
    a.view = pk.view
    a.B, a.H, a.W = domain
    a.ntaps = len(pk.taps)
    for i, (dy, dx, p) in enumerate(pk.taps):
        a.tap_dy[i], a.tap_dx[i], a.tap_p[i] = dy, dx, p
    a.weight = w.data_ptr()
    a.N, a.N_pad, a.K_pad = pk.n, pk.n_pad, pk.k_pad
    ptr = lambda t: None if t is None else t.data_ptr()
    sc = kw.get("shortcut")          # (sources, res_conv weight [N, C, 1, 1], split, bias): fused 1x1 shortcut
    sc_ref = None
    if sc is not None:
        from diffusion_models_b200.packing import append_shortcut, pack_conv
        w = dev(append_shortcut(pk, sc[1], sc[2]))
        a.weight, a.K_pad = w.data_ptr(), w.shape[1]
        a.rsrc0, a.rC0, a.rld0 = sc[0][0].data_ptr(), sc[0][0].shape[-1], sc[0][0].shape[-1]
        if len(sc[0]) > 1:
            a.rsrc1, a.rC1, a.rld1 = sc[0][1].data_ptr(), sc[0][1].shape[-1], sc[0][1].shape[-1]
        a.rbias = ptr(sc[3])
        wr = dev(pack_conv(sc[1], sc[2]).weight).float()[:pk.n]          # bf16-rounded, like the kernel's operand
        sc_ref = torch.cat([t.float() for t in sc[0]], dim=-1) @ wr.t() + sc[3]
    a.row_scale, a.bias, a.norm_g = ptr(kw.get("row_scale")), ptr(kw.get("bias")), ptr(kw.get("norm_g"))
    ss = kw.get("scale_shift")
    a.scale_shift = ptr(ss)
    a.ss_stride = ss.shape[1] if (ss is not None and ss.shape[0] > 1) else 0
    a.act = kw.get("act", 0)
    res = kw.get("residual")
    a.residual, a.ld_res = ptr(res), (res.shape[-1] if res is not None else 0)
    nchw = kw.get("out_f32_nchw", False)
    a.out, a.out_f32_nchw = out.data_ptr(), int(nchw)
    if nchw:
        a.ld_out, a.OH, a.OW = 0, out.shape[2], out.shape[3]
    else:
        a.ld_out, a.OH, a.OW = out.shape[-1], out.shape[1], out.shape[2]
    a.sy, a.sx, a.oy, a.ox = kw.get("out_map", (1, 1, 0, 0))
    rn = kw.get("rnorm_out")
    a.rnorm_out = ptr(rn)
    check(lib.ddm_conv2d(C.byref(a), stream()))
    ref_out = torch.zeros_like(out, dtype=F32)
    ref_rn = torch.zeros_like(rn) if rn is not None else None
    R.conv_ref([s.float() for s in srcs], dev(pk.weight).float(), pk.n, domain, pk.taps, view=pk.view,
               row_scale=kw.get("row_scale"), bias=kw.get("bias"), norm_g=kw.get("norm_g"), scale_shift=ss,
               act=a.act, residual=res.float() if res is not None else sc_ref, out=ref_out, out_map=(a.sy, a.sx, a.oy, a.ox),
               out_f32_nchw=nchw, rnorm_out=ref_rn)
    return ref_out, ref_rn


# ------------------------------------------------------------------------------------------------ conv kernel
def test_conv_wide_output_1024(lib):
    """C_out = 1024 (Unet(dim=128, dim_mults=(1,2,4,8)) bottleneck): four N tiles of 256."""
    from diffusion_models_b200.packing import pack_conv
    x = dev(rnd((4, 4, 4, 256), 301), BF)
    pk = pack_conv(rnd((1024, 256, 3, 3), 302, 0.02))
    bias = dev(rnd((1024,), 303, 0.1))
    out = torch.zeros((4, 4, 4, 1024), dtype=BF, device="cuda")
    ref, _ = run_conv(lib, [x], pk, (4, 4, 4), out, bias=bias)
    close(out, ref)


def test_gemm_1x1_plain(lib):
    from diffusion_models_b200.packing import pack_conv
    x = dev(rnd((2, 16, 16, 64), 1), BF)
    pk = pack_conv(rnd((64, 64, 1, 1), 2, 0.125))
    out = torch.zeros((2, 16, 16, 64), dtype=BF, device="cuda")
    ref, _ = run_conv(lib, [x], pk, (2, 16, 16), out)
    close(out, ref)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 32, 64, 64), (3, 16, 16, 128, 128), (4, 8, 8, 256, 256),
                                             (9, 4, 4, 256, 256), (1, 64, 64, 64, 64), (1, 8, 24, 32, 32)])
def test_conv3x3_block_epilogue(lib, B, H, W, Cin, Cout):
    """Block.forward: conv -> RMSNorm -> scale/shift -> SiLU (+ residual, + row-norm side output)."""
    from diffusion_models_b200.packing import pack_conv
    x = dev(rnd((B, H, W, Cin), 3), BF)
    pk = pack_conv(rnd((Cout, Cin, 3, 3), 4, (Cin * 9) ** -0.5))
    bias, g = dev(rnd((Cout,), 5, 0.1)), dev(1 + 0.1 * rnd((Cout,), 6)) * Cout ** 0.5
    ss = dev(rnd((B, 2 * Cout), 7, 0.3))
    res = dev(rnd((B, H, W, Cout), 8), BF)
    out = torch.zeros((B, H, W, Cout), dtype=BF, device="cuda")
    rn = torch.zeros((B * H * W,), dtype=F32, device="cuda")
    ref, ref_rn = run_conv(lib, [x], pk, (B, H, W), out, bias=bias, norm_g=g, scale_shift=ss, act=1, residual=res, rnorm_out=rn)
    close(out, ref)
    close(rn, ref_rn, 2e-2)


@pytest.mark.parametrize("B,H,W,Cin,Cout,with_res,with_norm", [
    (5, 32, 32, 64, 64, False, True),      # dx-folded kernel (C_out = 64, tile = whole rows, no residual), 2-issuer tile mode
    (3, 16, 16, 64, 64, False, True),      # folded, two image rows per warp
    (2, 10, 32, 64, 64, False, True),      # folded, ragged tile rows (H % 4 != 0)
    (2, 32, 32, 64, 64, False, False),     # folded, plain conv (no norm / activation)
    (2, 10, 32, 64, 64, True, True),       # lean kernel, residual tile by TMA, ragged rows
    (3, 32, 32, 128, 64, True, True),      # 2 chunks per tap, residual
    (1, 8, 64, 64, 64, True, True),        # 64 wide: two 32-wide tiles per row, taps cross the tile edge
    (2, 6, 48, 64, 64, True, True),        # 48 wide: second tile half outside the image, ragged rows
    (1, 4, 160, 32, 64, False, True),      # 160 wide, C_in = 32 (half-filled K chunk)
    (3, 32, 32, 128, 64, False, True),     # 128 -> 64 folded with resident weights (144 KB: the plan without alignment slack)
    (2, 10, 32, 128, 64, False, True),     # the same, ragged tile rows
    (160, 32, 32, 128, 64, False, True),   # the same, nine tiles per CTA: the two-stage slab ring wraps many times
])
def test_conv3x3_lean_kernel_variants(lib, B, H, W, Cin, Cout, with_res, with_norm):
    """The lean epilogue kernel's variants (dx-folded / residual by TMA / 32-wide tiles) with the batch-shared
    scale/shift row the sampler uses (Block.forward dd:113-122, ResnetBlock dd:136-148)."""
    from diffusion_models_b200.packing import pack_conv
    x = dev(rnd((B, H, W, Cin), 23), BF)
    pk = pack_conv(rnd((Cout, Cin, 3, 3), 24, (Cin * 9) ** -0.5))
    kw = dict(bias=dev(rnd((Cout,), 25, 0.1)))
    if with_norm:
        kw.update(norm_g=dev(1 + 0.1 * rnd((Cout,), 26)) * Cout ** 0.5, scale_shift=dev(rnd((1, 2 * Cout), 27, 0.3)), act=1)
    if with_res:
        kw.update(residual=dev(rnd((B, H, W, Cout), 28), BF))
    out = torch.zeros((B, H, W, Cout), dtype=BF, device="cuda")
    ref, _ = run_conv(lib, [x], pk, (B, H, W), out, **kw)
    close(out, ref)
    # the same launch twice gives the same bits (two MMA issuer threads must not make the sums order-dependent)
    out2 = torch.zeros_like(out)
    run_conv(lib, [x], pk, (B, H, W), out2, **kw)
    assert torch.equal(out, out2)


def test_conv3x3_shared_scale_shift_row(lib):
    from diffusion_models_b200.packing import pack_conv
    B, H, W, Cin, Cout = 2, 16, 16, 64, 64
    x = dev(rnd((B, H, W, Cin), 13), BF)
    pk = pack_conv(rnd((Cout, Cin, 3, 3), 14, (Cin * 9) ** -0.5))
    ss = dev(rnd((1, 2 * Cout), 17, 0.3))
    out = torch.zeros((B, H, W, Cout), dtype=BF, device="cuda")
    ref, _ = run_conv(lib, [x], pk, (B, H, W), out, bias=dev(rnd((Cout,), 15, 0.1)), norm_g=dev(torch.full((Cout,), 8.0)),
                      scale_shift=ss, act=1)
    close(out, ref)


def test_conv3x3_two_sources_concat(lib):
    from diffusion_models_b200.packing import pack_conv
    B, H, W = 2, 16, 16
    a, b = dev(rnd((B, H, W, 128), 20), BF), dev(rnd((B, H, W, 64), 21), BF)
    pk = pack_conv(rnd((128, 192, 3, 3), 22, (192 * 9) ** -0.5), split=(128, 64))
    out = torch.zeros((B, H, W, 128), dtype=BF, device="cuda")
    ref, _ = run_conv(lib, [a, b], pk, (B, H, W), out, bias=dev(rnd((128,), 23, 0.1)))
    close(out, ref)


def test_conv3x3_two_sources_folded_resident(lib):
    """ups.3.x.block1 of the benchmark net: cat(x, skip) 64 + 64 -> 64 at 32 x 32, Block epilogue, no residual."""
    from diffusion_models_b200.packing import pack_conv
    B, H, W = 40, 32, 32
    a, b = dev(rnd((B, H, W, 64), 220), BF), dev(rnd((B, H, W, 64), 221), BF)
    pk = pack_conv(rnd((64, 128, 3, 3), 222, (128 * 9) ** -0.5), split=(64, 64))
    out = torch.zeros((B, H, W, 64), dtype=BF, device="cuda")
    kw = dict(bias=dev(rnd((64,), 223, 0.1)), norm_g=dev(1 + 0.1 * rnd((64,), 224)) * 8.0, scale_shift=dev(rnd((1, 128), 225, 0.3)), act=1)
    ref, _ = run_conv(lib, [a, b], pk, (B, H, W), out, **kw)
    close(out, ref)
    out2 = torch.zeros_like(out)
    run_conv(lib, [a, b], pk, (B, H, W), out2, **kw)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("B,H,W,rc,N", [(3, 32, 32, (64, 64), 64), (150, 32, 32, (64, 64), 64), (2, 64, 64, (128,), 64), (1, 8, 32, (64, 64), 64),
                                        (5, 16, 16, (128, 64), 128), (300, 16, 16, (128, 64), 128), (2, 32, 32, (128,), 128)])
def test_conv3x3_fused_shortcut(lib, B, H, W, rc, N):
    """ResnetBlock tail with dim_in != dim_out (dd:134,148): block2's conv + RMSNorm + SiLU with res_conv(cat(x, skip)) + bias riding
    along as extra K steps into a second accumulator."""
    from diffusion_models_b200.packing import pack_conv
    assert lib.ddm_conv2d_shortcut_supported(N, N, rc[0], rc[1] if len(rc) > 1 else 0, H, W) == 1
    h1 = dev(rnd((B, H, W, N), 240), BF)
    rs = [dev(rnd((B, H, W, c), 241 + i), BF) for i, c in enumerate(rc)]
    pk = pack_conv(rnd((N, N, 3, 3), 244, (N * 9) ** -0.5))
    wr, br = rnd((N, sum(rc), 1, 1), 245, sum(rc) ** -0.5), dev(rnd((N,), 246, 0.1))
    kw = dict(bias=dev(rnd((N,), 247, 0.1)), norm_g=dev(1 + 0.1 * rnd((N,), 248)) * N ** 0.5, act=1, shortcut=(rs, wr, rc if len(rc) > 1 else None, br))
    out = torch.zeros((B, H, W, N), dtype=BF, device="cuda")
    ref, _ = run_conv(lib, [h1], pk, (B, H, W), out, **kw)
    close(out, ref)
    out2 = torch.zeros_like(out)
    run_conv(lib, [h1], pk, (B, H, W), out2, **kw)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("B,H,W,Cin,Cout,ks,with_norm", [(16, 4, 4, 512, 512, 4, True), (16, 4, 4, 256, 256, 9, True), (3, 4, 4, 256, 512, 12, True),
                                                         (32, 8, 8, 128, 256, 2, True), (5, 4, 4, 512, 256, 7, False), (2, 6, 10, 64, 128, 3, True)])
def test_conv_split_k(lib, B, H, W, Cin, Cout, ks, with_norm):
    """Few output rows (4x4 / 8x8 levels at small batches): ddm_conv2d with ksplit K ranges -> fp32 partial sums, and the Block tail
    (bias, RMSNorm, scale/shift, SiLU, residual) over their sum in ddm_rmsnorm_act_split."""
    from diffusion_models_b200._lib import ConvArgs
    from diffusion_models_b200.packing import pack_conv
    x = dev(rnd((B, H, W, Cin), 260), BF)
    pk = pack_conv(rnd((Cout, Cin, 3, 3), 261, (Cin * 9) ** -0.5))
    w = dev(pk.weight)
    rows = B * H * W
    part = torch.full((ks, rows, Cout), float("nan"), dtype=F32, device="cuda")       # every range must be written
    a = ConvArgs()
    a.src0, a.C0, a.ld0 = x.data_ptr(), Cin, Cin
    a.B, a.H, a.W = B, H, W
    a.ntaps = 9
    for i, (dy, dx, p) in enumerate(pk.taps):
        a.tap_dy[i], a.tap_dx[i], a.tap_p[i] = dy, dx, p
    a.weight, a.N, a.N_pad, a.K_pad = w.data_ptr(), Cout, pk.n_pad, pk.k_pad
    a.OH, a.OW, a.sy, a.sx, a.ld_out = H, W, 1, 1, Cout
    a.ksplit, a.partial_out = ks, part.data_ptr()
    check(lib.ddm_conv2d(C.byref(a), stream()))
    raw = torch.zeros((B, H, W, Cout), dtype=F32, device="cuda")
    R.conv_ref([x.float()], w.float(), Cout, (B, H, W), pk.taps, out=raw, round_out=False)
    got = part.sum(0).reshape(B, H, W, Cout)
    assert torch.isfinite(got).all()
    assert (got - raw).abs().max().item() <= 2e-3 * raw.abs().max().item()
    bias, g = dev(rnd((Cout,), 262, 0.1)), (dev(1 + 0.1 * rnd((Cout,), 263)) * Cout ** 0.5 if with_norm else None)
    ss = dev(rnd((1, 2 * Cout), 264, 0.3)) if with_norm else None
    res = dev(rnd((B, H, W, Cout), 265), BF)
    out = torch.zeros((B, H, W, Cout), dtype=BF, device="cuda")
    check(lib.ddm_rmsnorm_act_split(part.data_ptr(), ks, bias.data_ptr(), g.data_ptr() if with_norm else None, ss.data_ptr() if with_norm else None,
                                    0, H * W, 1 if with_norm else 0, res.data_ptr(), out.data_ptr(), rows, Cout, stream()))
    ref = torch.zeros((B, H, W, Cout), dtype=F32, device="cuda")
    R.conv_ref([x.float()], w.float(), Cout, (B, H, W), pk.taps, bias=bias, norm_g=g, scale_shift=ss, act=1 if with_norm else 0,
               residual=res.float(), out=ref)
    close(out, ref)
    sug = lib.ddm_conv2d_suggest_ksplit(rows, pk.n_pad, pk.k_pad)
    assert 1 <= sug <= 16 and (sug - 1) * -(-(pk.k_pad // 64) // sug) < pk.k_pad // 64


@pytest.mark.parametrize("B,H,W,Cin,Cout,batched", [(8, 4, 4, 256, 512, False), (37, 4, 4, 512, 512, False), (300, 4, 4, 128, 512, False),
                                                      (5, 8, 8, 64, 384, True), (2, 6, 10, 128, 512, False)])
def test_conv_row_norm_in_cta_pair(lib, B, H, W, Cin, Cout, batched):
    """Block epilogue on rows wider than one accumulator (C_out = 384 / 512): the two N tiles of a row in a CTA pair, sums of
    squares exchanged through distributed shared memory.  37 / 300 images: 5 / 38 M tiles, i.e. clusters that walk several tiles
    and an odd tile count; 6 x 10: ragged tiles."""
    from diffusion_models_b200.packing import pack_conv
    assert lib.ddm_conv2d_row_norm_supported(Cout) == 1
    x = dev(rnd((B, H, W, Cin), 280), BF)
    pk = pack_conv(rnd((Cout, Cin, 3, 3), 281, (Cin * 9) ** -0.5))
    kw = dict(bias=dev(rnd((Cout,), 282, 0.1)), norm_g=dev(1 + 0.1 * rnd((Cout,), 283)) * Cout ** 0.5,
              scale_shift=dev(rnd((B if batched else 1, 2 * Cout), 284, 0.3)), act=1, residual=dev(rnd((B, H, W, Cout), 285), BF))
    out = torch.zeros((B, H, W, Cout), dtype=BF, device="cuda")
    ref, _ = run_conv(lib, [x], pk, (B, H, W), out, **kw)
    close(out, ref)
    out2 = torch.zeros_like(out)
    run_conv(lib, [x], pk, (B, H, W), out2, **kw)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("B,H,W,N,nh,with_sc", [(3, 32, 32, 64, 3, True), (150, 32, 32, 64, 3, True), (2, 10, 24, 64, 1, False), (3, 16, 16, 128, 4, True),
                                                 (2, 32, 32, 64, 2, False)])
def test_conv3x3_fused_head(lib, B, H, W, N, nh, with_sc):
    """final_res_block.block2 + final_conv (dd:343,390): the 1x1 head applied to the tile's fp32 values in the epilogue, fp32 NCHW out;
    with the fused shortcut (the benchmark net) or a plain residual; ragged tiles (10 x 24)."""
    from diffusion_models_b200._lib import ConvArgs
    from diffusion_models_b200.packing import append_shortcut, pack_conv
    assert lib.ddm_conv2d_head_supported(N, nh, H, W) == 1
    h1 = dev(rnd((B, H, W, N), 300), BF)
    pk = pack_conv(rnd((N, N, 3, 3), 301, (N * 9) ** -0.5))
    bias, g = dev(rnd((N,), 302, 0.1)), dev(1 + 0.1 * rnd((N,), 303)) * N ** 0.5
    hw, hb = dev(rnd((nh, N), 304, N ** -0.5)), dev(rnd((nh,), 305, 0.1))
    out = torch.full((B, nh, H, W), float("nan"), dtype=F32, device="cuda")
    a = ConvArgs()
    a.src0, a.C0, a.ld0 = h1.data_ptr(), N, N
    a.B, a.H, a.W, a.ntaps = B, H, W, 9
    for i, (dy, dx, p) in enumerate(pk.taps):
        a.tap_dy[i], a.tap_dx[i], a.tap_p[i] = dy, dx, p
    w = dev(pk.weight)
    if with_sc:
        rs = [dev(rnd((B, H, W, 64), 306 + i), BF) for i in range(2)]
        wr, br = rnd((N, 128, 1, 1), 308, 128 ** -0.5), dev(rnd((N,), 309, 0.1))
        w = dev(append_shortcut(pk, wr, (64, 64)))
        a.rsrc0, a.rC0, a.rld0, a.rsrc1, a.rC1, a.rld1, a.rbias = rs[0].data_ptr(), 64, 64, rs[1].data_ptr(), 64, 64, br.data_ptr()
        res = torch.cat([t.float() for t in rs], dim=-1) @ dev(pack_conv(wr, (64, 64)).weight).float()[:N].t() + br
    else:
        rt = dev(rnd((B, H, W, N), 310), BF)
        a.residual, a.ld_res = rt.data_ptr(), N
        res = rt.float()
    a.weight, a.N, a.N_pad, a.K_pad = w.data_ptr(), N, pk.n_pad, w.shape[1]
    a.bias, a.norm_g, a.act = bias.data_ptr(), g.data_ptr(), 1
    a.ld_out, a.OH, a.OW, a.sy, a.sx = N, H, W, 1, 1
    a.head_w, a.head_b, a.head_out, a.head_n = hw.data_ptr(), hb.data_ptr(), out.data_ptr(), nh
    check(lib.ddm_conv2d(C.byref(a), stream()))
    v = torch.zeros((B, H, W, N), dtype=F32, device="cuda")
    R.conv_ref([h1.float()], dev(pk.weight).float(), N, (B, H, W), pk.taps, bias=bias, norm_g=g, act=1, residual=res, out=v, round_out=False)
    ref = (v.reshape(-1, N) @ hw.t() + hb).reshape(B, H, W, nh).permute(0, 3, 1, 2)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    out2 = torch.zeros_like(out)
    a.head_out = out2.data_ptr()
    check(lib.ddm_conv2d(C.byref(a), stream()))
    assert torch.equal(out, out2)


def test_conv_wide_output_two_n_tiles(lib):
    """N = 384 (to_qkv, two 192-wide tiles, pre-norm row scale) and N = 512 (two 256-wide tiles)."""
    from diffusion_models_b200.packing import pack_conv
    B, H, W = 2, 8, 8
    x = dev(rnd((B, H, W, 128), 30), BF)
    rs = dev(torch.rand((B * H * W,), generator=torch.Generator().manual_seed(31)) + 0.5)
    pk = pack_conv(rnd((384, 128, 1, 1), 32, 128 ** -0.5))
    out = torch.zeros((B, H, W, 384), dtype=BF, device="cuda")
    ref, _ = run_conv(lib, [x], pk, (B, H, W), out, row_scale=rs)
    close(out, ref)
    x2 = dev(rnd((B, 4, 4, 256), 33), BF)
    pk2 = pack_conv(rnd((512, 256, 3, 3), 34, (256 * 9) ** -0.5))
    out2 = torch.zeros((B, 4, 4, 512), dtype=BF, device="cuda")
    ref2, _ = run_conv(lib, [x2], pk2, (B, 4, 4), out2, bias=dev(rnd((512,), 35, 0.1)))
    close(out2, ref2)


def test_downsample_unshuffle_view(lib):
    from diffusion_models_b200.packing import pack_downsample
    B, H, W, C_, Co = 2, 16, 16, 64, 128           # source is [B,32,32,64]
    x = dev(rnd((B, 2 * H, 2 * W, C_), 40), BF)
    pk = pack_downsample(rnd((Co, 4 * C_, 1, 1), 41, (4 * C_) ** -0.5))
    out = torch.zeros((B, H, W, Co), dtype=BF, device="cuda")
    ref, _ = run_conv(lib, [x], pk, (B, H, W), out, bias=dev(rnd((Co,), 42, 0.1)))
    close(out, ref)
    # and against the reference formulation itself: Rearrange + 1x1 conv (dd:54-58)
    wt = rnd((Co, 4 * C_, 1, 1), 41, (4 * C_) ** -0.5).to(BF).float().cuda()
    xn = x.float().permute(0, 3, 1, 2)
    un = xn.reshape(B, C_, H, 2, W, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, 4 * C_, H, W)
    y = torch.nn.functional.conv2d(un, wt, dev(rnd((Co,), 42, 0.1)))
    close(out, y.permute(0, 2, 3, 1))


def test_upsample_four_phases(lib):
    from diffusion_models_b200.packing import pack_upsample
    B, H, W, C_, Co = 2, 8, 8, 128, 64
    x = dev(rnd((B, H, W, C_), 50), BF)
    wt = rnd((Co, C_, 3, 3), 51, (C_ * 9) ** -0.5)
    bias = dev(rnd((Co,), 52, 0.1))
    out = torch.zeros((B, 2 * H, 2 * W, Co), dtype=BF, device="cuda")
    for pk, ph, pw in pack_upsample(wt):
        run_conv(lib, [x], pk, (B, H, W), out, bias=bias, out_map=(2, 2, ph, pw))
    up = x.float().permute(0, 3, 1, 2).repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    y = torch.nn.functional.conv2d(up, wt.cuda(), bias, padding=1)           # dd:48-52
    close(out, y.permute(0, 2, 3, 1), 2e-2)


def test_final_conv_fp32_nchw(lib):
    from diffusion_models_b200.packing import pack_conv
    B, H, W = 3, 32, 32
    x = dev(rnd((B, H, W, 64), 60), BF)
    pk = pack_conv(rnd((3, 64, 1, 1), 61, 0.125))
    out = torch.zeros((B, 3, H, W), dtype=F32, device="cuda")
    ref, _ = run_conv(lib, [x], pk, (B, H, W), out, bias=dev(rnd((3,), 62, 0.1)), out_f32_nchw=True)
    close(out, ref, 2e-3)


def test_token_gemm_ragged_rows(lib):
    """nn.Linear over a [1,1,rows,K] token matrix whose row count is not a multiple of the 128-row tile."""
    from diffusion_models_b200.packing import pack_linear
    rows = 5 * 77
    x = dev(rnd((1, 1, rows, 512), 70), BF)
    pk = pack_linear(rnd((128, 512), 71, 512 ** -0.5))
    out = torch.zeros((1, 1, rows, 128), dtype=BF, device="cuda")
    ref, _ = run_conv(lib, [x], pk, (1, 1, rows), out)
    close(out, ref)


def test_conv_persistent_many_tiles(lib):
    """More tiles than SMs: exercises the persistent loop, smem ring wrap-around and both TMEM accumulator stages."""
    from diffusion_models_b200.packing import pack_conv
    B, H, W = 40, 32, 32          # 320 tiles
    x = dev(rnd((B, H, W, 64), 80), BF)
    pk = pack_conv(rnd((64, 64, 3, 3), 81, (64 * 9) ** -0.5))
    out = torch.zeros((B, H, W, 64), dtype=BF, device="cuda")
    ref, _ = run_conv(lib, [x], pk, (B, H, W), out, bias=dev(rnd((64,), 82, 0.1)), norm_g=dev(torch.full((64,), 8.0)), act=1)
    close(out, ref)


# ------------------------------------------------------------------------------------------------ other kernels
@pytest.mark.parametrize("cs,B,H,W,Co", [((3, 0, 0), 2, 32, 40, 64), ((4, 4, 0), 2, 32, 40, 64), ((3, 3, 2), 2, 32, 40, 64),
                                         ((3, 0, 0), 96, 32, 40, 64), ((3, 0, 0), 300, 32, 32, 64), ((2, 2, 0), 3, 10, 24, 32),
                                         ((1, 0, 0), 2, 7, 9, 128), ((3, 0, 0), 2, 64, 64, 64), ((1, 1, 1), 5, 33, 70, 64)])
def test_stem_conv(lib, cs, B, H, W, Co):
    """C_in <= 4: tcgen05 over an explicit im2col tile; more channels: mma.sync.  B = 96 / 300: 1440 / 2400 tiles > the persistent
    grid, so every CTA walks several tiles with the next patch prefetched; 10 x 24, 7 x 9, 33 x 70: ragged tiles."""
    from diffusion_models_b200.packing import pack_stem
    ins = [dev(rnd((B, c, H, W), 90 + i)) for i, c in enumerate(cs) if c]
    cin = sum(cs)
    wt = rnd((Co, cin, 7, 7), 95, (cin * 49) ** -0.5)
    wp, bias = dev(pack_stem(wt)), dev(rnd((Co,), 96, 0.1))
    out = torch.zeros((B, H, W, Co), dtype=BF, device="cuda")
    p = [t.data_ptr() for t in ins] + [None] * (3 - len(ins))
    c = [x for x in cs if x] + [0] * (3 - len(ins))
    check(lib.ddm_stem_conv(p[0], c[0], p[1], c[1], p[2], c[2], wp.data_ptr(), bias.data_ptr(), out.data_ptr(), B, H, W, Co, 7, stream()))
    close(out, R.stem_ref(ins, wp, bias, 7, Co))


def test_time_path_kernels(lib):
    t = dev(torch.tensor([0.0, 17.0, 999.0, 500.0]))
    emb = torch.zeros((4, 64), device="cuda")
    check(lib.ddm_sinusoidal_embedding(t.data_ptr(), emb.data_ptr(), 4, 64, 10000.0, stream()))
    assert (emb - R.sinusoidal_ref(t, 64, 10000.0)).abs().max().item() < 2e-4
    x, W_, b = dev(rnd((4, 256), 100)), dev(rnd((300, 256), 101, 256 ** -0.5)), dev(rnd((300,), 102, 0.1))
    for ai, ao in ((0, 0), (1, 0), (0, 2)):
        y = torch.zeros((4, 300), device="cuda")
        check(lib.ddm_small_linear(x.data_ptr(), 256, W_.data_ptr(), b.data_ptr(), y.data_ptr(), 300, 4, 300, 256, ai, ao, stream()))
        assert (y - R.small_linear_ref(x, W_, b, ai, ao)).abs().max().item() < 1e-4


@pytest.mark.parametrize("C_", [32, 64, 128, 256, 512])
def test_row_norm_kernels(lib, C_):
    rows = 1000
    x = dev(rnd((rows, C_), 110), BF)
    rn = torch.zeros((rows,), device="cuda")
    check(lib.ddm_row_rnorm(x.data_ptr(), C_, rn.data_ptr(), rows, C_, stream()))
    close(rn, R.row_rnorm_ref(x), 1e-4)
    g, ss, res = dev(rnd((C_,), 111) * 0.1 + 1) * C_ ** 0.5, dev(rnd((4, 2 * C_), 112, 0.3)), dev(rnd((rows, C_), 113), BF)
    out = torch.zeros((rows, C_), dtype=BF, device="cuda")
    check(lib.ddm_rmsnorm_act(x.data_ptr(), g.data_ptr(), ss.data_ptr(), 2 * C_, 250, 1, res.data_ptr(), out.data_ptr(), rows, C_, stream()))
    close(out, R.rmsnorm_act_ref(x.float(), g, ss, 250, 1, res.float()))


@pytest.mark.parametrize("n,heads,d", [(1024, 4, 32), (64, 4, 32), (256, 2, 16), (100, 4, 64), (4096, 4, 32), (16384, 4, 32)])
def test_linear_attention(lib, n, heads, d):
    B = 3 if n <= 1024 else 2          # 4096 / 16384 tokens: the 64-px / 128-px levels of BASELINE configs 3-5
    qkv = dev(rnd((B, n, 3 * heads * d), 120), BF)
    mem = dev(rnd((2, heads, d, 4), 121))
    out = torch.zeros((B, n, heads * d), dtype=BF, device="cuda")
    check(lib.ddm_linear_attention(qkv.data_ptr(), mem.data_ptr(), out.data_ptr(), B, n, heads, d, 4, stream()))
    close(out, R.linear_attention_ref(qkv.float(), mem, heads, d), 2e-2)


@pytest.mark.parametrize("n,slack", [(64, 0.0), (64, 6.0), (1024, 2.0), (100, 0.5)])
def test_linear_attention_bounded_shift(lib, n, slack):
    """ddm_linear_attention_bounded: any per-channel upper bound of k (here: the true maximum + slack, also >= the memory keys) gives
    the same softmax over the tokens as the max pass."""
    B, heads, d = 3, 4, 32
    qkv = dev(rnd((B, n, 3 * heads * d), 122), BF)
    mem = dev(rnd((2, heads, d, 4), 123))
    k = qkv.float()[:, :, heads * d:2 * heads * d]
    shift = (torch.maximum(k.amax(dim=(0, 1)), mem[0].reshape(heads * d, -1).amax(dim=1)) + slack).contiguous()
    out = torch.zeros((B, n, heads * d), dtype=BF, device="cuda")
    check(lib.ddm_linear_attention_bounded(qkv.data_ptr(), mem.data_ptr(), shift.data_ptr(), out.data_ptr(), B, n, heads, d, 4, stream()))
    close(out, R.linear_attention_ref(qkv.float(), mem, heads, d), 2e-2)


@pytest.mark.parametrize("C_", [64, 128])
@pytest.mark.parametrize("B,n,wscale,n_mem", [(3, 1024, 1.0, 4), (2, 128, 1.0, 4), (5, 256, 3.0, 4), (300, 256, 1.0, 4), (2, 4096, 1.0, 4),
                                              (1, 16384, 0.5, 2), (3, 384, 1.0, 0)])
def test_linear_attention_block_fused(lib, B, n, wscale, n_mem, C_):
    """ddm_linear_attention_block (one tcgen05 kernel) against the unfused fp32 statement of dd:173-193 + residual."""
    from diffusion_models_b200._lib import LinAttnBlockArgs
    from diffusion_models_b200.packing import linattn_k_shift, norm_gain, pack_conv
    heads, d = 4, 32
    hid = heads * d
    assert lib.ddm_linear_attention_block_supported(C_, n, heads, d, n_mem) == 1
    x = dev(rnd((B, n, C_), 200) * (1 + rnd((B, n, 1), 201).abs()), BF)      # rows of different lengths
    w_qkv = rnd((3 * hid, C_, 1, 1), 202, wscale / C_ ** 0.5)
    g_in = rnd((1, C_, 1, 1), 203) * 0.1 + 1
    w_out = rnd((C_, hid, 1, 1), 204, 0.09)
    b_out = dev(rnd((C_,), 205, 0.1))
    g_out = dev(norm_gain(rnd((1, C_, 1, 1), 206) * 0.1 + 1))
    mem = rnd((2, heads, d, max(n_mem, 1)), 207)[..., :n_mem].contiguous()
    pq, po = dev(pack_conv(w_qkv, in_scale=norm_gain(g_in)).weight), dev(pack_conv(w_out).weight)
    shift = dev(linattn_k_shift(w_qkv, g_in, mem, heads, d))
    memd = dev(mem) if n_mem else None
    out = torch.zeros((B, n, C_), dtype=BF, device="cuda")
    a = LinAttnBlockArgs()
    a.x, a.out, a.B, a.n, a.C = x.data_ptr(), out.data_ptr(), B, n, C_
    a.w_qkv, a.w_out, a.bias_out, a.g_out = pq.data_ptr(), po.data_ptr(), b_out.data_ptr(), g_out.data_ptr()
    a.mem_kv, a.k_shift = (memd.data_ptr() if n_mem else None), shift.data_ptr()
    a.heads, a.dim_head, a.n_mem = heads, d, n_mem
    check(lib.ddm_linear_attention_block(C.byref(a), stream()))
    ref = R.linattn_block_ref(x.float(), pq.float(), po.float(), b_out, g_out, dev(mem), heads, d)
    close(out, ref, 2e-2)
    out2 = torch.zeros_like(out)                   # bitwise repeatable, and a second launch on the same buffers is clean
    a.out = out2.data_ptr()
    check(lib.ddm_linear_attention_block(C.byref(a), stream()))
    assert torch.equal(out, out2)


@pytest.mark.parametrize("nq,nk,heads,d,n_mem", [(16, 16, 4, 32, 4), (64, 64, 4, 32, 4), (256, 256, 4, 32, 4), (64, 77, 4, 32, 0),
                                                 (16, 1, 4, 32, 0), (16, 16, 2, 16, 4), (256, 256, 1, 128, 0), (100, 300, 2, 64, 3),
                                                 (1024, 1024, 1, 128, 0), (16, 16, 4, 32, 16)])
def test_softmax_attention(lib, nq, nk, heads, d, n_mem):
    """ddm_attention: tcgen05 kernel (d = 32 / 64 / 128; tiles spanning several images, several key tiles, memory keys) and the
    CUDA-core kernel (d = 16)."""
    B, hd = (3 if nq <= 256 else 2), heads * d
    if nq == 16 and nk == 16 and n_mem == 4 and d == 32:
        B = 37                  # 592 query rows: 4 full tiles of 8 images + a ragged one
    q, k, v = (dev(rnd((B, n_, hd), 130 + i), BF) for i, n_ in enumerate((nq, nk, nk)))
    mk, mv = (dev(rnd((heads, max(n_mem, 1), d), 135 + i)) for i in range(2))
    out = torch.zeros((B, nq, hd), dtype=BF, device="cuda")
    check(lib.ddm_attention(q.data_ptr(), hd, k.data_ptr(), hd, v.data_ptr(), hd, mk.data_ptr() if n_mem else None,
                            mv.data_ptr() if n_mem else None, n_mem, out.data_ptr(), B, nq, nk, heads, d, stream()))
    close(out, R.attention_ref(q.float(), k.float(), v.float(), mk if n_mem else None, mv if n_mem else None, heads, d), 2e-2)


@pytest.mark.parametrize("kind,objective", [(0, 0), (0, 1), (0, 2), (1, 0), (1, 2)])
def test_sampler_step_bit_exact(lib, kind, objective):
    """fp32 update must match the reference's unfused op chain bit for bit (no FMA contraction)."""
    shape = (2, 3, 32, 32)
    x, mo, z = dev(rnd(shape, 140)), dev(rnd(shape, 141)), dev(rnd((3,) + shape, 142))
    coef = torch.tensor([[1.9, 1.6, 0.7, 0.6, 0.3, 0.0, 0.52, 0.85],
                         [1.2, 0.66, 0.9, 0.4, 0.0, 0.0, 0.83, 0.55],
                         [1.01, 0.14, 0.0, 0.0, 0.0, 1.0, 0.99, 0.14]], device="cuda")
    counter = torch.zeros((1,), dtype=torch.int32, device="cuda")
    xs, x0 = x.clone(), torch.zeros_like(x)
    ref = x.clone()
    for s in range(3):
        check(lib.ddm_sampler_step(kind, xs.data_ptr(), mo.data_ptr(), z.data_ptr(), x.numel(), x0.data_ptr(), coef.data_ptr(),
                                   counter.data_ptr(), 1, objective, 0, x.numel(), stream()))
        ref, ref0 = R.sampler_step_ref(kind, ref, mo, z[s], coef[s], objective)
        assert torch.equal(xs, ref), (s, (xs - ref).abs().max().item())
        assert torch.equal(x0, ref0)
    assert counter.item() == 3


def test_sampler_step_learned_variance(lib):
    """LearnedGaussianDiffusion ancestral step (lgd:91-111, dd:638-645) against the same chain of torch fp32 ops."""
    B, C_, H, W = 2, 3, 16, 16
    shape = (B, C_, H, W)
    x, mo, z = dev(rnd(shape, 150)), dev(rnd((B, 2 * C_, H, W), 151)), dev(rnd((3,) + shape, 152))
    coef = torch.tensor([[1.9, 1.6, 0.7, 0.6, 1.0, -4.0, -1.5, 0.0],
                         [1.2, 0.66, 0.9, 0.4, 1.0, -6.0, -3.0, 0.0],
                         [1.01, 0.14, 1.0, 0.0, 0.0, -46.0, -9.0, 0.0]], device="cuda")
    counter = torch.zeros((2,), dtype=torch.int32, device="cuda")
    xs, x0 = x.clone(), torch.zeros_like(x)
    ref = x.clone()
    for s in range(3):
        check(lib.ddm_sampler_step_learned(xs.data_ptr(), mo.data_ptr(), z.data_ptr(), x.numel(), x0.data_ptr(), coef.data_ptr(),
                                           counter.data_ptr(), 1, 0, x.numel(), C_ * H * W, stream()))
        ra, rm1, c1, c2, on, lo, hi = (coef[s, i] for i in range(7))
        eps, v = mo.chunk(2, dim=1)
        frac = (v + 1) * 0.5
        logvar = frac * hi + (1 - frac) * lo
        r0 = (ra * ref - rm1 * eps).clamp(-1.0, 1.0)
        mean = c1 * r0 + c2 * ref
        ref = mean + (0.5 * logvar).exp() * z[s] if on.item() != 0 else mean
        assert torch.equal(x0, r0)
        assert (xs - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())     # exp() may differ in the last ulp
    assert counter[0].item() == 3


def test_philox_normal_statistics(lib):
    n = 1 << 20
    x = torch.zeros((n,), device="cuda")
    check(lib.ddm_randn(x.data_ptr(), 1234, 0, n, stream()))
    y = torch.zeros((n,), device="cuda")
    check(lib.ddm_randn(y.data_ptr(), 1234, 1, n, stream()))
    assert abs(x.mean().item()) < 5e-3 and abs(x.std().item() - 1) < 5e-3
    assert abs((x * y).mean().item()) < 5e-3            # different streams are uncorrelated
    assert abs((x ** 4).mean().item() - 3.0) < 0.05      # kurtosis of N(0,1)
    out = torch.zeros((n,), device="cuda")
    check(lib.ddm_finalize(x.data_ptr(), out.data_ptr(), 1, n, stream()))
    assert torch.equal(out, (x + 1) * 0.5)


@pytest.mark.parametrize("N,C_", [(3, 64), (4, 64), (8, 32), (6, 128)])
def test_head_conv1x1(lib, N, C_):
    """final_conv: bf16 channels-last -> fp32 NCHW with fp32 weights (dd:343,390)."""
    B, H, W = 3, 16, 24
    x = dev(rnd((B, H, W, C_), 150), BF)
    w, b = dev(rnd((N, C_), 151, C_ ** -0.5)), dev(rnd((N,), 152, 0.1))
    out = torch.zeros((B, N, H, W), dtype=F32, device="cuda")
    check(lib.ddm_head_conv1x1(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, H * W, C_, N, stream()))
    # fp32 reference on the CPU (cuDNN would use TF32 by default)
    ref = torch.einsum("bhwc,nc->bnhw", x.float().cpu(), w.cpu()) + b.cpu()[None, :, None, None]
    close(out, ref, 1e-4)
