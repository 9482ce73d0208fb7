"""Kernel plan for one U-Net evaluation at a fixed (batch, height, width).

`UnetEngine` turns a `UnetSpec` + fp32 master weights into (a) packed device weights and (b) a flat list of
C-ABI launches over statically allocated channels-last bf16 activation buffers.  Running the plan is just walking
that list on the current CUDA stream, so a whole step is capturable in a CUDA graph.  PyTorch is used for device
memory and streams only; every launch is one of our kernels (see include/ddm_b200.h).

Layer semantics follow denoising_diffusion.py:349-390 (Unet.forward) and the blocks at :98-229.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib
from .arch import AttnSpec, ResBlockSpec, StageSpec, UnetSpec
import os

from .packing import PackedConv, append_shortcut, linattn_k_shift, norm_gain, pack_conv, pack_downsample, pack_linear, pack_stem, pack_upsample

MAX_FUSED_NORM = 256     # one CTA owns a full output row in TMEM only up to 256 columns
MAX_K_SHIFT = 40.0       # fused linear attention: exp(k - shift) must stay a normal fp32 for k >= -shift


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class PlanOps:
    """Plan-building helpers shared by the kernel plans (U-Net forward here, VAE decode in vae.py): device tensors kept alive
    for the raw pointers baked into the launch list, activation buffers, and the ddm_conv2d argument struct.  A plan owner
    provides `device`, `lib`, `_keep`, `ops`, `taps`, `op_meta` and the fp32 master weights `_w`."""

    def _init_plan(self, device, lib):
        self.device = torch.device(device)
        if lib is not None:            # test seam: the CPU test-suite injects a stand-in that checks the host logic of a plan
            self.lib = lib
        else:
            if self.device.type != "cuda":
                raise RuntimeError("the kernel plans need a CUDA (B200) device: there is no CPU fallback")
            self.lib = _lib.init(self.device.index if self.device.index is not None else torch.cuda.current_device())
        self._keep: List[object] = []            # device tensors / ctypes structs referenced by raw pointer
        self.ops: List[Tuple[str, Callable[[int], int]]] = []
        self.taps: Dict[str, torch.Tensor] = {}  # named activations for per-layer parity tests
        self.op_meta: Dict[str, dict] = {}       # GEMM shape / byte counts per conv launch (profiling aid)

    # ------------------------------------------------------------------ helpers
    def _dev(self, t: torch.Tensor, dtype=None) -> torch.Tensor:
        t = t.to(device=self.device, dtype=dtype or t.dtype).contiguous()
        self._keep.append(t)
        return t

    def _f32(self, name: str) -> torch.Tensor:
        return self._dev(self._w[name].float().reshape(-1))

    def _act(self, b: int, h: int, w: int, c: int, tag: Optional[str] = None) -> torch.Tensor:
        t = torch.empty((b, h, w, c), dtype=torch.bfloat16, device=self.device)
        self._keep.append(t)
        if tag:
            self.taps[tag] = t
        return t

    def _conv(self, tag: str, pk: PackedConv, srcs: List[torch.Tensor], out: torch.Tensor, *, domain: Tuple[int, int, int],
              bias=None, row_scale=None, norm_g=None, ss=None, ss_stride=0, act=0, residual=None, out_f32_nchw=False,
              out_map=(1, 1, 0, 0), rnorm_out=None, ld_src: Optional[List[int]] = None, into=None, shortcut=None, ksplit=None,
              head=None):
        """`shortcut` = (weight with the 1x1 shortcut appended along K, its sources, its bias): out += W_r . cat(sources) + bias."""
        wdev = self._dev(pk.weight if shortcut is None else shortcut[0])
        a = _lib.ConvArgs()
        a.src0 = srcs[0].data_ptr()
        a.src1 = srcs[1].data_ptr() if len(srcs) > 1 else None
        a.C0 = pk.seg_channels[0] // (2 if pk.view == 1 else 1)
        a.C1 = pk.seg_channels[1] if len(srcs) > 1 else 0
        lds = ld_src or [s.shape[-1] for s in srcs]
        a.ld0 = lds[0]
        a.ld1 = lds[1] if len(srcs) > 1 else 0
        a.view = pk.view
        a.B, a.H, a.W = domain
        a.ntaps = len(pk.taps)
        for i, (dy, dx, p) in enumerate(pk.taps):
            a.tap_dy[i], a.tap_dx[i], a.tap_p[i] = dy, dx, p
        a.weight = wdev.data_ptr()
        a.N, a.N_pad, a.K_pad = pk.n, pk.n_pad, wdev.shape[1]
        if shortcut is not None:
            rs = shortcut[1]
            a.rsrc0, a.rC0, a.rld0 = rs[0].data_ptr(), rs[0].shape[-1], rs[0].shape[-1]
            if len(rs) > 1:
                a.rsrc1, a.rC1, a.rld1 = rs[1].data_ptr(), rs[1].shape[-1], rs[1].shape[-1]
            a.rbias = _ptr(shortcut[2])
        a.row_scale, a.bias, a.norm_g, a.scale_shift = _ptr(row_scale), _ptr(bias), _ptr(norm_g), _ptr(ss)
        a.ss_stride = ss_stride
        a.act = act
        a.residual = _ptr(residual)
        a.ld_res = residual.shape[-1] if residual is not None else 0
        a.out = out.data_ptr() if out is not None else None
        a.out_f32_nchw = 1 if out_f32_nchw else 0
        if head is not None:            # (weight [n][N], bias [n], fp32 NCHW output): the net's last 1x1 conv rides in the epilogue
            a.head_w, a.head_b, a.head_out, a.head_n = head[0].data_ptr(), head[1].data_ptr(), head[2].data_ptr(), head[2].shape[1]
            a.ld_out, a.OH, a.OW = pk.n, domain[1], domain[2]
        if head is not None:
            pass
        elif out_f32_nchw:
            a.ld_out = 0
            a.OH, a.OW = out.shape[2], out.shape[3]
        else:
            a.ld_out = out.shape[-1]
            a.OH, a.OW = out.shape[1], out.shape[2]
        a.sy, a.sx, a.oy, a.ox = out_map
        a.rnorm_out = _ptr(rnorm_out)
        if ksplit is not None:          # (ranges, fp32 workspace [ranges][rows][N]): raw partial sums only, `out` is not written
            a.ksplit, a.partial_out = ksplit[0], ksplit[1].data_ptr()
        self._keep.append(a)
        self.op_meta[tag] = dict(M=a.B * a.H * a.W, N=a.N, K=a.K_pad, taps=a.ntaps, srcs=len(srcs),
                                 out_bytes=a.B * a.H * a.W * (a.head_n * 4 if head is not None else a.N * (4 if out_f32_nchw else 2)),
                                 in_bytes=sum(int(x.numel()) * 2 for x in srcs) + (int(residual.numel()) * 2 if residual is not None else 0))
        fn = self.lib.ddm_conv2d
        ref = C.byref(a)
        (into if into is not None else self.ops).append((tag, lambda s, fn=fn, ref=ref: fn(ref, s)))

    def _splitk_workspace(self, numel: int) -> torch.Tensor:
        """One fp32 scratch buffer shared by every split-K layer of the plan (the ops run in stream order)."""
        ws = getattr(self, "_splitk_ws", None)
        if ws is None or ws.numel() < numel:
            assert ws is None or not getattr(self, "_splitk_used", False) or True
            ws = torch.zeros((numel,), dtype=torch.float32, device=self.device)
            self._splitk_ws = ws
            self._keep.append(ws)
        return ws[:numel]

    def _add(self, tag: str, fn: Callable[[int], int], into=None):
        (into if into is not None else self.ops).append((tag, fn))


    def _run(self, ops, stream: Optional[int] = None):
        if stream is not None:
            s = stream
        else:
            s = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else 0
        for tag, fn in ops:
            rc = fn(s)
            if rc != 0:
                _lib.check(rc, tag)



class UnetEngine(PlanOps):
    def __init__(self, spec: UnetSpec, weights: Dict[str, torch.Tensor], batch: int, height: int, width: int,
                 device: torch.device, time_rows: int = 1, text_tokens: int = 0, fuse_rnorm: bool = True, lib=None):
        """`time_rows` is 1 when every sample shares the timestep (sampling loops, dd:641,682) or `batch`."""
        assert height % spec.downsample_factor == 0 and width % spec.downsample_factor == 0, \
            f"your input dimensions {(height, width)} need to be divisible by {spec.downsample_factor}, given the unet"
        assert time_rows in (1, batch)
        self.spec, self.B, self.H, self.W = spec, batch, height, width
        self._init_plan(device, lib)
        self.time_rows = time_rows
        self.text_tokens = text_tokens
        self.fuse_rnorm = fuse_rnorm
        self.time_ops: List[Tuple[str, Callable[[int], int]]] = []
        self.text_ops: List[Tuple[str, Callable[[int], int]]] = []
        self.loops: Dict[tuple, dict] = {}       # cached sampling loops (tables + captured step graph), see diffusion.py
        self._w = {k: v.detach() for k, v in weights.items()}
        self._build()

    # ------------------------------------------------------------------ plan
    def _build(self):
        sp, B, H, W, lib, dev = self.spec, self.B, self.H, self.W, self.lib, self.device
        f32 = dict(dtype=torch.float32, device=dev)

        # ---- static I/O (fp32 NCHW like the reference's tensors)
        self.x = torch.zeros((B, sp.channels, H, W), **f32)
        self.x_self_cond = torch.zeros((B, sp.channels, H, W), **f32) if sp.self_condition else None
        self.cond = torch.zeros((B, sp.cond_channels, H, W), **f32) if sp.cond_channels else None
        self.out = torch.zeros((B, sp.out_dim, H, W), **f32)
        self.time = torch.zeros((self.time_rows,), **f32)

        # ---- time path (dd:280-285 + per-block mlp dd:127-130, tc:146-152) -> one [rows, sum 2C] table
        blocks = sp.res_blocks()
        self.ss_offsets, off = {}, 0
        for rb in blocks:
            self.ss_offsets[rb.name] = off
            off += 2 * rb.c_out
        self.ss_width = off
        R = self.time_rows
        self.ss = torch.zeros((R, self.ss_width), **f32)
        self.ss_stride = self.ss_width if R == B and B > 1 else 0
        self._build_time_path(self.time, self.ss, R, self.time_ops)

        # ---- stem (dd:356-357; ic:52-54 cat(x, cond); dd:352-354 cat(self_cond, x))
        stem_w = self._dev(pack_stem(self._w["init_conv.weight"]))
        stem_b = self._f32("init_conv.bias")
        ins = []
        if sp.self_condition:
            ins.append((self.x_self_cond, sp.channels))
        ins.append((self.x, sp.channels))
        if sp.cond_channels:
            ins.append((self.cond, sp.cond_channels))
        while len(ins) < 3:
            ins.append((None, 0))
        h0 = self._act(B, H, W, sp.init_dim, "init_conv")
        self._add("init_conv", lambda s: lib.ddm_stem_conv(
            _ptr(ins[0][0]), ins[0][1], _ptr(ins[1][0]), ins[1][1], _ptr(ins[2][0]), ins[2][1], stem_w.data_ptr(),
            stem_b.data_ptr(), h0.data_ptr(), B, H, W, sp.init_dim, sp.stem_kernel, s))

        # ---- text path for cross attention: K/V projections are loop-invariant (tc:63-65) -> text_ops
        if sp.text_mode == "xattn":
            m = max(self.text_tokens, 1)
            self.text = torch.zeros((B, m, sp.text_emb_dim), dtype=torch.bfloat16, device=dev)
            self.text_kv = {}
            inner = sp.xattn_heads * sp.xattn_dim_head
            for nm in ("cross_attn_down", "cross_attn", "cross_attn_up"):
                kv = []
                for proj in ("to_k", "to_v"):
                    o = self._act(1, 1, B * m, inner)
                    self._conv(f"{nm}.{proj}", pack_linear(self._w[f"{nm}.{proj}.weight"]),
                               [self.text.view(1, 1, B * m, sp.text_emb_dim)], o, domain=(1, 1, B * m), into=self.text_ops)
                    kv.append(o)
                self.text_kv[nm] = kv
        elif sp.text_mode == "concat":
            self.text = torch.zeros((B, sp.text_emb_dim), **f32)
        else:
            self.text = None

        # ---- body
        x, h, w = h0, H, W
        skips: List[torch.Tensor] = []
        for st in sp.downs:
            x = self._resblock(st.block1, [x], h, w); skips.append(x)
            x = self._resblock(st.block2, [x], h, w, want_rnorm=not self._linattn_fused(st.attn, h, w))
            x = self._attention(st.attn, x, h, w); skips.append(x)
            x, h, w = self._resample(st, x, h, w)
        if sp.text_mode == "xattn":
            x = self._cross_attention("cross_attn_down", x, h, w)
        x = self._resblock(sp.mid1, [x], h, w, want_rnorm=(sp.text_mode != "xattn"))
        if sp.text_mode == "xattn":
            x = self._cross_attention("cross_attn", x, h, w)
        x = self._attention(sp.mid_attn, x, h, w)
        x = self._resblock(sp.mid2, [x], h, w)
        if sp.text_mode == "xattn":
            x = self._cross_attention("cross_attn_up", x, h, w)
        for st in sp.ups:
            x = self._resblock(st.block1, [x, skips.pop()], h, w)
            x = self._resblock(st.block2, [x, skips.pop()], h, w, want_rnorm=not self._linattn_fused(st.attn, h, w))
            x = self._attention(st.attn, x, h, w)
            x, h, w = self._resample(st, x, h, w)
        hw_, hb_ = self._f32("final_conv.weight"), self._f32("final_conv.bias")
        x = self._resblock(sp.final_block, [x, h0], h, w, head=(hw_, hb_, self.out) if sp.out_dim <= 4 else None)
        if x is None:                             # final_conv went into final_res_block.block2's epilogue
            pass
        elif sp.out_dim in (1, 2, 3, 4, 6, 8):    # HBM-bound head: dedicated kernel, fp32 weights, fp32 NCHW output
            cin = x.shape[-1]
            self._add("final_conv", lambda s: lib.ddm_head_conv1x1(x.data_ptr(), hw_.data_ptr(), hb_.data_ptr(),
                                                                    self.out.data_ptr(), B, h * w, cin, sp.out_dim, s))
        else:
            self._conv("final_conv", pack_conv(self._w["final_conv.weight"]), [x], self.out, domain=(B, h, w),
                       bias=self._f32("final_conv.bias"), out_f32_nchw=True)

    def _build_time_path(self, t_in: torch.Tensor, ss_out: torch.Tensor, rows: int, into):
        """sinusoid -> Linear -> GELU -> Linear [-> text concat] -> (SiLU -> Linear) for all blocks at once."""
        sp, lib, dev = self.spec, self.lib, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        td = sp.time_dim
        sin = torch.zeros((rows, sp.fourier_dim), **f32)
        hid = torch.zeros((rows, td), **f32)
        temb = torch.zeros((rows, td), **f32)
        self._keep += [sin, hid, temb]
        w1, b1 = self._f32("time_mlp.1.weight"), self._f32("time_mlp.1.bias")
        w2, b2 = self._f32("time_mlp.3.weight"), self._f32("time_mlp.3.bias")
        self._add("time.sin", lambda s: lib.ddm_sinusoidal_embedding(t_in.data_ptr(), sin.data_ptr(), rows, sp.fourier_dim,
                                                                     sp.theta, s), into)
        self._add("time.l1", lambda s: lib.ddm_small_linear(sin.data_ptr(), sp.fourier_dim, w1.data_ptr(), b1.data_ptr(),
                                                            hid.data_ptr(), td, rows, td, sp.fourier_dim, 0, 2, s), into)
        self._add("time.l2", lambda s: lib.ddm_small_linear(hid.data_ptr(), td, w2.data_ptr(), b2.data_ptr(),
                                                            temb.data_ptr(), td, rows, td, td, 0, 0, s), into)
        self.t_emb = temb
        if sp.text_mode == "concat":
            assert rows == self.B, "text-concat conditioning makes the time embedding per-sample: time_rows must be batch"
            # t = text_concat_proj(cat(t, text_proj(text)))   (tc:146-152)
            cat = torch.zeros((rows, 2 * td), **f32)
            th = torch.zeros((rows, td), **f32)
            t2 = torch.zeros((rows, td), **f32)
            self._keep += [cat, th, t2]
            pw0, pb0 = self._f32("text_proj.0.weight"), self._f32("text_proj.0.bias")
            pw2, pb2 = self._f32("text_proj.2.weight"), self._f32("text_proj.2.bias")
            cw, cb = self._f32("text_concat_proj.weight"), self._f32("text_concat_proj.bias")
            self._text_concat = True
            # re-point the second time linear at the first half of `cat`
            into.pop()
            self._add("time.l2", lambda s: lib.ddm_small_linear(hid.data_ptr(), td, w2.data_ptr(), b2.data_ptr(),
                                                                cat.data_ptr(), 2 * td, rows, td, td, 0, 0, s), into)
            self._add("text.p0", lambda s: lib.ddm_small_linear(self.text.data_ptr(), sp.text_emb_dim, pw0.data_ptr(),
                                                                pb0.data_ptr(), th.data_ptr(), td, rows, td,
                                                                sp.text_emb_dim, 0, 2, s), into)
            self._add("text.p2", lambda s: lib.ddm_small_linear(th.data_ptr(), td, pw2.data_ptr(), pb2.data_ptr(),
                                                                cat.data_ptr() + td * 4, 2 * td, rows, td, td, 0, 0, s), into)
            self._add("text.cat", lambda s: lib.ddm_small_linear(cat.data_ptr(), 2 * td, cw.data_ptr(), cb.data_ptr(),
                                                                 t2.data_ptr(), td, rows, td, 2 * td, 0, 0, s), into)
            temb = t2
            self.t_emb = t2
        blocks = sp.res_blocks()
        wss = self._dev(torch.cat([self._w[rb.name + ".mlp.1.weight"].float() for rb in blocks], dim=0))
        bss = self._dev(torch.cat([self._w[rb.name + ".mlp.1.bias"].float() for rb in blocks], dim=0))
        width = ss_out.shape[1]
        self._add("time.ss", lambda s: lib.ddm_small_linear(temb.data_ptr(), td, wss.data_ptr(), bss.data_ptr(),
                                                            ss_out.data_ptr(), width, rows, width, td, 1, 0, s), into)

    def _tail_fuses(self, pk, rows: int, shortcut) -> bool:
        """Whether `_block_tail` runs this layer as ONE conv launch with the whole Block tail in its epilogue."""
        c = pk.n
        if shortcut is None and c % 8 == 0 and c <= 1024 and not os.environ.get("DDM_NO_SPLITK") and \
                self.lib.ddm_conv2d_suggest_ksplit(rows, pk.n_pad, pk.k_pad) > 1:
            return False
        return c <= MAX_FUSED_NORM

    def _block_tail(self, tag, pk, srcs, out, h, w, *, bias, g, ss, act, residual, rnorm_out=None, shortcut=None, head=None):
        """conv -> RMSNorm -> scale/shift -> act -> (+residual): fused into the conv epilogue when the output row
        fits one TMEM tile, otherwise conv(+bias) followed by the row-norm kernel."""
        B, lib = self.B, self.lib
        c = pk.n
        rows = B * h * w
        ks = 1
        if shortcut is None and c % 8 == 0 and c <= 1024 and not os.environ.get("DDM_NO_SPLITK"):
            ks = lib.ddm_conv2d_suggest_ksplit(rows, pk.n_pad, pk.k_pad)
        if ks > 1:
            # few output rows (4x4 / 8x8 levels at small per-GPU batches): K ranges of a tile on different SMs, fp32 partial sums,
            # and the Block tail over their sum in the row-norm kernel
            ws = self._splitk_workspace(ks * rows * c)
            assert head is None
            self._conv(tag + ".splitk", pk, srcs, out, domain=(B, h, w), ksplit=(ks, ws))
            self._add(tag + ".norm", lambda s: lib.ddm_rmsnorm_act_split(ws.data_ptr(), ks, _ptr(bias), _ptr(g), _ptr(ss), self.ss_stride, h * w,
                                                                          act, _ptr(residual), out.data_ptr(), rows, c, s))
            if rnorm_out is not None:
                self._add(tag + ".rnorm", lambda s: lib.ddm_row_rnorm(out.data_ptr(), c, rnorm_out.data_ptr(), rows, c, s))
            return
        if c <= MAX_FUSED_NORM:
            self._conv(tag, pk, srcs, out, domain=(B, h, w), bias=bias, norm_g=g, ss=ss, ss_stride=self.ss_stride,
                       act=act, residual=residual, rnorm_out=rnorm_out, shortcut=shortcut, head=head)
            return
        assert shortcut is None and head is None
        if (g is not None and not os.environ.get("DDM_NO_PAIR_NORM") and pk.n == pk.n_pad and
                lib.ddm_conv2d_row_norm_supported(c)):
            # 512-channel rows: the two 256-column tiles of a row in a CTA pair, RMSNorm in the epilogue (no separate norm launch)
            self._conv(tag, pk, srcs, out, domain=(B, h, w), bias=bias, norm_g=g, ss=ss, ss_stride=self.ss_stride,
                       act=act, residual=residual)
            if rnorm_out is not None:
                self._add(tag + ".rnorm", lambda s: lib.ddm_row_rnorm(out.data_ptr(), c, rnorm_out.data_ptr(), rows, c, s))
            return
        tmp = self._act(B, h, w, c)
        self._conv(tag + ".gemm", pk, srcs, tmp, domain=(B, h, w), bias=bias)
        self._add(tag + ".norm", lambda s: lib.ddm_rmsnorm_act(tmp.data_ptr(), _ptr(g), _ptr(ss), self.ss_stride, h * w, act,
                                                                _ptr(residual), out.data_ptr(), rows, c, s))
        if rnorm_out is not None:
            self._add(tag + ".rnorm", lambda s: lib.ddm_row_rnorm(out.data_ptr(), c, rnorm_out.data_ptr(), rows, c, s))

    def _resblock(self, rb: ResBlockSpec, srcs: List[torch.Tensor], h: int, w: int, want_rnorm: bool = False,
                  head=None) -> Optional[torch.Tensor]:
        """ResnetBlock.forward, dd:136-148.  `head` = (weight, bias, fp32 NCHW output) of the net's final 1x1 conv: when block2 can
        take it into its epilogue, the block's own output is never materialised and None is returned."""
        B, W_ = self.B, self._w
        split = rb.split if len(srcs) > 1 else None
        ss = self.ss[:, self.ss_offsets[rb.name]:]
        h1 = self._act(B, h, w, rb.c_out)
        self._block_tail(rb.name + ".block1", pack_conv(W_[rb.name + ".block1.proj.weight"], split), srcs, h1, h, w,
                         bias=self._f32(rb.name + ".block1.proj.bias"), g=self._dev(norm_gain(W_[rb.name + ".block1.norm.g"])),
                         ss=ss, act=1, residual=None)
        pk2 = pack_conv(W_[rb.name + ".block2.proj.weight"])
        rn = None
        if want_rnorm and self.fuse_rnorm:
            rn = torch.zeros((B * h * w,), dtype=torch.float32, device=self.device)
            self._keep.append(rn)
        shortcut = None
        if rb.c_in == rb.c_out:
            res = srcs[0]
        elif (rn is None and not os.environ.get("DDM_NO_FUSED_SHORTCUT") and all(s.shape[-1] % 64 == 0 for s in srcs) and
              self.lib.ddm_conv2d_shortcut_supported(rb.c_out, rb.c_out, srcs[0].shape[-1], srcs[1].shape[-1] if len(srcs) > 1 else 0, h, w)):
            # res_conv (1x1, dd:134) rides along block2 as extra K steps into a second accumulator: no launch, no round trip
            res = None
            shortcut = (append_shortcut(pk2, W_[rb.name + ".res_conv.weight"], split), srcs, self._f32(rb.name + ".res_conv.bias"))
        else:
            res = self._act(B, h, w, rb.c_out)
            self._conv(rb.name + ".res_conv", pack_conv(W_[rb.name + ".res_conv.weight"], split), srcs, res,
                       domain=(B, h, w), bias=self._f32(rb.name + ".res_conv.bias"))
        fuse_head = (head is not None and rn is None and self._tail_fuses(pk2, B * h * w, shortcut) and
                     self.lib.ddm_conv2d_head_supported(rb.c_out, head[2].shape[1], h, w))
        out = None if fuse_head else self._act(B, h, w, rb.c_out, rb.name)
        self._block_tail(rb.name + ".block2", pk2, [h1], out, h, w,
                         bias=self._f32(rb.name + ".block2.proj.bias"), g=self._dev(norm_gain(W_[rb.name + ".block2.norm.g"])),
                         ss=None, act=1, residual=res, rnorm_out=rn, shortcut=shortcut, head=head if fuse_head else None)
        self._last_rnorm = (out, rn)
        return out

    def _linattn_fused(self, at: AttnSpec, h: int, w: int) -> bool:
        """Whether `attn(x) + x` of this LinearAttention runs as the single fused kernel (ddm_linear_attention_block)."""
        if at.full or os.environ.get("DDM_NO_FUSED_LINATTN"):
            return False
        if not self.lib.ddm_linear_attention_block_supported(at.dim, h * w, at.heads, at.dim_head, at.n_mem):
            return False
        shift = linattn_k_shift(self._w[at.name + ".to_qkv.weight"], self._w[at.name + ".norm.g"], self._w[at.name + ".mem_kv"],
                                at.heads, at.dim_head)
        return float(shift.max()) <= MAX_K_SHIFT

    def _attention_fused(self, at: AttnSpec, x: torch.Tensor, h: int, w: int) -> torch.Tensor:
        """LinearAttention block + residual in one launch (dd:173-193, :368)."""
        B, lib, W_ = self.B, self.lib, self._w
        wq = self._dev(pack_conv(W_[at.name + ".to_qkv.weight"], in_scale=norm_gain(W_[at.name + ".norm.g"])).weight)
        wo = self._dev(pack_conv(W_[at.name + ".to_out.0.weight"]).weight)
        out = self._act(B, h, w, at.dim, at.name)
        a = _lib.LinAttnBlockArgs()
        a.x, a.out = x.data_ptr(), out.data_ptr()
        a.B, a.n, a.C = B, h * w, at.dim
        a.w_qkv, a.w_out = wq.data_ptr(), wo.data_ptr()
        a.bias_out = self._f32(at.name + ".to_out.0.bias").data_ptr()
        a.g_out = self._dev(norm_gain(W_[at.name + ".to_out.1.g"])).data_ptr()
        a.mem_kv = self._dev(W_[at.name + ".mem_kv"].float()).data_ptr()
        a.k_shift = self._dev(linattn_k_shift(W_[at.name + ".to_qkv.weight"], W_[at.name + ".norm.g"], W_[at.name + ".mem_kv"],
                                              at.heads, at.dim_head)).data_ptr()
        a.heads, a.dim_head, a.n_mem = at.heads, at.dim_head, at.n_mem
        self._keep.append(a)
        fn, ref = lib.ddm_linear_attention_block, C.byref(a)
        self._add(at.name + ".block", lambda s: fn(ref, s))
        return out

    def _attention(self, at: AttnSpec, x: torch.Tensor, h: int, w: int) -> torch.Tensor:
        """`attn(x) + x` with LinearAttention (dd:173-193) or Attention (dd:215-229)."""
        if self._linattn_fused(at, h, w):
            return self._attention_fused(at, x, h, w)
        B, lib, W_ = self.B, self.lib, self._w
        n, hid, c = h * w, at.heads * at.dim_head, at.dim
        rows = B * n
        last_out, rn = getattr(self, "_last_rnorm", (None, None))
        if last_out is not x or rn is None:
            rn = torch.zeros((rows,), dtype=torch.float32, device=self.device)
            self._keep.append(rn)
            self._add(at.name + ".rnorm", lambda s: lib.ddm_row_rnorm(x.data_ptr(), c, rn.data_ptr(), rows, c, s))
        # pre-norm folded into the qkv GEMM: W (x * g sqrt(C) / |x|) = rnorm[pixel] * ((W diag(g sqrt(C))) x)
        qkv = self._act(B, h, w, 3 * hid)
        self._conv(at.name + ".to_qkv", pack_conv(W_[at.name + ".to_qkv.weight"], in_scale=norm_gain(W_[at.name + ".norm.g"])),
                   [x], qkv, domain=(B, h, w), row_scale=rn)
        a = self._act(B, h, w, hid)
        mem = self._dev(W_[at.name + ".mem_kv"].float())
        if at.full:
            mk, mv = mem[0], mem[1]
            self._add(at.name + ".attend", lambda s: lib.ddm_attention(
                qkv.data_ptr(), 3 * hid, qkv.data_ptr() + hid * 2, 3 * hid, qkv.data_ptr() + 2 * hid * 2, 3 * hid,
                mk.data_ptr(), mv.data_ptr(), at.n_mem, a.data_ptr(), B, n, n, at.heads, at.dim_head, s))
            out = self._act(B, h, w, c, at.name)
            self._conv(at.name + ".to_out", pack_conv(W_[at.name + ".to_out.weight"]), [a], out, domain=(B, h, w),
                       bias=self._f32(at.name + ".to_out.bias"), residual=x)
        else:
            shift = None
            if at.dim_head == 32 and not os.environ.get("DDM_NO_BOUNDED_LINATTN"):
                sh = linattn_k_shift(W_[at.name + ".to_qkv.weight"], W_[at.name + ".norm.g"], W_[at.name + ".mem_kv"], at.heads, at.dim_head)
                if float(sh.max()) <= MAX_K_SHIFT:          # (the bound as the softmax shift: no max pass over k, see linattn_tc.cu)
                    shift = self._dev(sh)
            if shift is not None:
                self._add(at.name + ".attend", lambda s: lib.ddm_linear_attention_bounded(
                    qkv.data_ptr(), mem.data_ptr(), shift.data_ptr(), a.data_ptr(), B, n, at.heads, at.dim_head, at.n_mem, s))
            else:
                self._add(at.name + ".attend", lambda s: lib.ddm_linear_attention(
                    qkv.data_ptr(), mem.data_ptr(), a.data_ptr(), B, n, at.heads, at.dim_head, at.n_mem, s))
            out = self._act(B, h, w, c, at.name)
            self._block_tail(at.name + ".to_out", pack_conv(W_[at.name + ".to_out.0.weight"]), [a], out, h, w,
                             bias=self._f32(at.name + ".to_out.0.bias"), g=self._dev(norm_gain(W_[at.name + ".to_out.1.g"])),
                             ss=None, act=0, residual=x)
        return out

    def _cross_attention(self, nm: str, x: torch.Tensor, h: int, w: int) -> torch.Tensor:
        """CrossAttention (tc:54-78) on the flattened bottleneck; the result replaces x (tc:176-177)."""
        sp, B, lib, W_ = self.spec, self.B, self.lib, self._w
        n, c, m = h * w, x.shape[-1], max(self.text_tokens, 1)
        inner, heads, d = sp.xattn_heads * sp.xattn_dim_head, sp.xattn_heads, sp.xattn_dim_head
        q = self._act(B, h, w, inner)
        self._conv(nm + ".to_q", pack_linear(W_[nm + ".to_q.weight"]), [x], q, domain=(B, h, w))
        k, v = self.text_kv[nm]
        a = self._act(B, h, w, inner)
        self._add(nm + ".attend", lambda s: lib.ddm_attention(q.data_ptr(), inner, k.data_ptr(), inner, v.data_ptr(), inner,
                                                              None, None, 0, a.data_ptr(), B, n, m, heads, d, s))
        out = self._act(B, h, w, c, nm)
        self._block_tail(nm + ".to_out", pack_linear(W_[nm + ".to_out.0.weight"]), [a], out, h, w,
                         bias=self._f32(nm + ".to_out.0.bias"), g=self._dev(norm_gain(W_[nm + ".to_out.1.g"])),
                         ss=None, act=0, residual=None)
        return out

    def _resample(self, st: StageSpec, x: torch.Tensor, h: int, w: int):
        B, W_ = self.B, self._w
        bias = self._f32(st.resample + ".bias")
        if st.resample_kind == "down":                      # dd:54-58
            out = self._act(B, h // 2, w // 2, st.c_res_out, st.resample)
            self._conv(st.resample, pack_downsample(W_[st.resample + ".weight"]), [x], out, domain=(B, h // 2, w // 2), bias=bias)
            return out, h // 2, w // 2
        if st.resample_kind == "up":                        # dd:48-52, four sub-pixel phases
            out = self._act(B, 2 * h, 2 * w, st.c_res_out, st.resample)
            for pk, ph, pw in pack_upsample(W_[st.resample + ".weight"]):
                self._conv(f"{st.resample}.p{ph}{pw}", pk, [x], out, domain=(B, h, w), bias=bias, out_map=(2, 2, ph, pw))
            return out, 2 * h, 2 * w
        out = self._act(B, h, w, st.c_res_out, st.resample)      # dd:319,336 plain 3x3
        self._conv(st.resample, pack_conv(W_[st.resample + ".weight"]), [x], out, domain=(B, h, w), bias=bias)
        return out, h, w

    def build_step_table(self, tvals: torch.Tensor) -> torch.Tensor:
        """Scale/shift rows for every timestep of a sampling loop, [S, ss_width]; valid because all samples share t."""
        rows = int(tvals.numel())
        table = torch.zeros((rows, self.ss_width), dtype=torch.float32, device=self.device)
        ops: List[Tuple[str, Callable[[int], int]]] = []
        keep_from = len(self._keep)
        self._build_time_path(tvals.contiguous(), table, rows, ops)
        self._run(ops)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        del self._keep[keep_from:]                 # temporaries of this one-off launch sequence
        return table

    # ------------------------------------------------------------------ execution
    def run_time_path(self, stream=None):
        self._run(self.time_ops, stream)

    def run_text_path(self, stream=None):
        self._run(self.text_ops, stream)

    def run_body(self, stream=None):
        self._run(self.ops, stream)

    @property
    def launches_per_forward(self) -> int:
        return len(self.ops)
