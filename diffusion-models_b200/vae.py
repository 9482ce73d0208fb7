"""VAE decode after the latent sampling loop on the B200 kernels.

`VQDecoder` stands in for the decode side of the reference's first-stage model -- `VQModel.decode` =
`post_quant_conv` -> `Decoder` (latent-diffusion/ldm/models/autoencoder.py:113-116, ldm/modules/diffusionmodules/model.py:
479-585) -- under the reference's own parameter names (`post_quant_conv.*`, `decoder.*`), so a `VQModel` checkpoint's
`state_dict` loads with `strict=False` (its encoder / quantiser / loss tensors are ignored).  Pass it as `vae=` to the
`LatentDiffusion` classes: `sample()` then runs the loop AND the decode on this package's kernels.

Plan of one decode (all launches through include/ddm_b200.h, nothing in PyTorch):
  post_quant_conv 1x1 (fp32 NCHW latents -> bf16 channels-last, padded to 8 channels)      ddm_stem_conv, k = 1
  every 3x3 / 1x1 conv, with bias and the block's residual add fused                      ddm_conv2d (tcgen05)
  GroupNorm(32, eps 1e-6) + swish in front of each conv                                   ddm_groupnorm_act
  AttnBlock: q | k | v as one 1x1 GEMM, single-head softmax attention (d = C)             ddm_conv2d + ddm_attention (tcgen05)
  nearest-2x Upsample + conv as four sub-pixel phases                                     ddm_conv2d, strided TMA stores
  conv_out straight to fp32 NCHW                                                          ddm_conv2d

GroupNorm statistics span a whole image, so the norm is a pre-pass of the conv, not part of the producing conv's epilogue.
Encoding (`VQModel.encode`, the quantiser) is not on the sampling path and stays with the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import torch
from torch import nn

from . import _lib
from .engine import PlanOps, _ptr
from .packing import pack_conv, pack_stem, pack_upsample
from .unet import _Holder, _init_param, _register          # parameter tree under dotted reference names

GROUPS, GN_EPS = 32, 1e-6


def decoder_param_shapes(*, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, resolution, z_channels, embed_dim,
                         attn_type="vanilla", **ignored) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """name -> (shape, init kind) of VQModel's decode side, in the reference's construction order (model.py:480-550)."""
    P: Dict[str, Tuple[Tuple[int, ...], str]] = {}

    def conv(name, co, ci, k):
        P[name + ".weight"] = ((co, ci, k, k), "conv")
        P[name + ".bias"] = ((co,), "bias:%d" % (ci * k * k))

    def norm(name, c):
        P[name + ".weight"] = ((c,), "ones")
        P[name + ".bias"] = ((c,), "zeros")

    def resblock(name, ci, co):
        norm(name + ".norm1", ci); conv(name + ".conv1", co, ci, 3)
        norm(name + ".norm2", co); conv(name + ".conv2", co, co, 3)
        if ci != co:
            conv(name + ".nin_shortcut", co, ci, 1)

    def attn(name, c):
        norm(name + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(f"{name}.{n}", c, c, 1)

    conv("post_quant_conv", z_channels, embed_dim, 1)
    nres = len(ch_mult)
    block_in = ch * ch_mult[nres - 1]
    curr_res = resolution // 2 ** (nres - 1)
    conv("decoder.conv_in", block_in, z_channels, 3)
    resblock("decoder.mid.block_1", block_in, block_in)
    if attn_type == "vanilla":
        attn("decoder.mid.attn_1", block_in)
    elif attn_type != "none":
        raise ValueError("only the reference's default attn_type='vanilla' (or 'none') is supported")
    resblock("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(nres)):
        block_out = ch * ch_mult[lvl]
        for j in range(num_res_blocks + 1):
            resblock(f"decoder.up.{lvl}.block.{j}", block_in, block_out)
            block_in = block_out
            if curr_res in attn_resolutions:
                attn(f"decoder.up.{lvl}.attn.{j}", block_in)
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
            curr_res *= 2
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", out_ch, block_in, 3)
    return P


class VaeDecodeEngine(PlanOps):
    """Kernel plan of one `decode` at a fixed (batch, latent height, latent width)."""

    def __init__(self, weights: Dict[str, torch.Tensor], batch: int, height: int, width: int, device, lib=None):
        self.B, self.h, self.w = batch, height, width
        self._init_plan(device, lib)
        self._w = {k: v.detach() for k, v in weights.items()}
        self._build()

    def _gn(self, tag: str, p: str, x: torch.Tensor, hw: int, act: int) -> torch.Tensor:
        lib, B, c = self.lib, self.B, x.shape[-1]
        if c % GROUPS != 0:
            raise ValueError(f"{p}: GroupNorm(32) needs a channel count divisible by 32, got {c}")
        out = self._act(B, x.shape[1], x.shape[2], c)
        g, b = self._f32(p + ".weight"), self._f32(p + ".bias")
        self._add(tag, lambda s: lib.ddm_groupnorm_act(x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), B, hw, c, GROUPS, GN_EPS, act, s))
        return out

    def _conv3(self, tag, p, x, h, w, residual=None, out=None, **kw):
        W_ = self._w
        co = W_[p + ".weight"].shape[0]
        out = out if out is not None else self._act(self.B, h, w, co, tag)
        self._conv(tag, pack_conv(W_[p + ".weight"]), [x], out, domain=(self.B, h, w), bias=self._f32(p + ".bias"), residual=residual, **kw)
        return out

    def _resblock(self, p: str, x: torch.Tensor, h: int, w: int) -> torch.Tensor:
        """model.py:115-138: conv1(swish(norm1 x)), conv2(swish(norm2 .)), + x (through nin_shortcut when widths differ)."""
        W_ = self._w
        a = self._gn(p + ".norm1", p + ".norm1", x, h * w, 1)
        h1 = self._conv3(p + ".conv1", p + ".conv1", a, h, w)
        b = self._gn(p + ".norm2", p + ".norm2", h1, h * w, 1)
        res = x
        if p + ".nin_shortcut.weight" in W_:
            res = self._conv3(p + ".nin_shortcut", p + ".nin_shortcut", x, h, w)
        elif p + ".conv_shortcut.weight" in W_:
            res = self._conv3(p + ".conv_shortcut", p + ".conv_shortcut", x, h, w)
        return self._conv3(p, p + ".conv2", b, h, w, residual=res)

    def _attn(self, p: str, x: torch.Tensor, h: int, w: int) -> torch.Tensor:
        """model.py:190-215."""
        W_, lib, B = self._w, self.lib, self.B
        c, n = x.shape[-1], h * w
        if c not in (32, 64, 128):
            raise ValueError(f"{p}: the tcgen05 attention kernel takes head dims 32 / 64 / 128; AttnBlock has one head of {c}")
        hn = self._gn(p + ".norm", p + ".norm", x, n, 0)
        wqkv = torch.cat([W_[f"{p}.{k}.weight"].float() for k in ("q", "k", "v")], dim=0)
        bqkv = self._dev(torch.cat([W_[f"{p}.{k}.bias"].float() for k in ("q", "k", "v")], dim=0))
        qkv = self._act(B, h, w, 3 * c)
        self._conv(p + ".qkv", pack_conv(wqkv), [hn], qkv, domain=(B, h, w), bias=bqkv)
        a = self._act(B, h, w, c)
        self._add(p + ".attend", lambda s: lib.ddm_attention(qkv.data_ptr(), 3 * c, qkv.data_ptr() + 2 * c, 3 * c, qkv.data_ptr() + 4 * c, 3 * c,
                                                              None, None, 0, a.data_ptr(), B, n, n, 1, c, s))
        return self._conv3(p, p + ".proj_out", a, h, w, residual=x)

    def _build(self):
        W_, B, h, w, lib, dev = self._w, self.B, self.h, self.w, self.lib, self.device
        zc_in = W_["post_quant_conv.weight"].shape[1]
        zc = W_["post_quant_conv.weight"].shape[0]
        self.z = torch.zeros((B, zc_in, h, w), dtype=torch.float32, device=dev)
        # post_quant_conv (autoencoder.py:114): 1x1 on the fp32 NCHW latents, written bf16 channels-last with the channel
        # count padded to 8 (zero weights / bias), which the tcgen05 conv reads as one 16-byte unit per pixel
        zpad = max(8, (zc + 7) // 8 * 8)
        wpq = torch.zeros((zpad, zc_in, 1, 1)); wpq[:zc] = W_["post_quant_conv.weight"].float().cpu()
        bpq = torch.zeros((zpad,)); bpq[:zc] = W_["post_quant_conv.bias"].float().cpu()
        wpq_d, bpq_d = self._dev(pack_stem(wpq)), self._dev(bpq)
        z8 = self._act(B, h, w, zpad, "post_quant_conv")
        self._add("post_quant_conv", lambda s, h=h, w=w: lib.ddm_stem_conv(self.z.data_ptr(), zc_in, None, 0, None, 0, wpq_d.data_ptr(),
                                                                           bpq_d.data_ptr(), z8.data_ptr(), B, h, w, zpad, 1, s))
        win = W_["decoder.conv_in.weight"].float().cpu()
        win8 = torch.zeros((win.shape[0], zpad, 3, 3)); win8[:, :zc] = win
        x = self._act(B, h, w, win.shape[0], "decoder.conv_in")
        self._conv("decoder.conv_in", pack_conv(win8), [z8], x, domain=(B, h, w), bias=self._f32("decoder.conv_in.bias"))
        x = self._resblock("decoder.mid.block_1", x, h, w)
        if "decoder.mid.attn_1.q.weight" in W_:
            x = self._attn("decoder.mid.attn_1", x, h, w)
        x = self._resblock("decoder.mid.block_2", x, h, w)
        levels = 0
        while f"decoder.up.{levels}.block.0.conv1.weight" in W_:
            levels += 1
        for lvl in reversed(range(levels)):
            j = 0
            while f"decoder.up.{lvl}.block.{j}.conv1.weight" in W_:
                x = self._resblock(f"decoder.up.{lvl}.block.{j}", x, h, w)
                if f"decoder.up.{lvl}.attn.{j}.q.weight" in W_:
                    x = self._attn(f"decoder.up.{lvl}.attn.{j}", x, h, w)
                j += 1
            up = f"decoder.up.{lvl}.upsample.conv"
            if up + ".weight" in W_:                     # model.py:70-73: nearest 2x then 3x3 conv = four sub-pixel phases
                c = W_[up + ".weight"].shape[0]
                out = self._act(B, 2 * h, 2 * w, c, up)
                bias = self._f32(up + ".bias")
                for pk, ph, pw in pack_upsample(W_[up + ".weight"]):
                    self._conv(f"{up}.p{ph}{pw}", pk, [x], out, domain=(B, h, w), bias=bias, out_map=(2, 2, ph, pw))
                x, h, w = out, 2 * h, 2 * w
        a = self._gn("decoder.norm_out", "decoder.norm_out", x, h * w, 1)
        oc = W_["decoder.conv_out.weight"].shape[0]
        self.out = torch.zeros((B, oc, h, w), dtype=torch.float32, device=dev)
        self._conv("decoder.conv_out", pack_conv(W_["decoder.conv_out.weight"]), [a], self.out, domain=(B, h, w),
                   bias=self._f32("decoder.conv_out.bias"), out_f32_nchw=True)

    def run(self, stream=None):
        self._run(self.ops, stream)


class VQDecoder(nn.Module):
    """Decode side of the reference's `VQModel` (autoencoder.py:19-116) with the same ddconfig keys and parameter names.

        vae = VQDecoder(ddconfig=dict(ch=64, out_ch=3, ch_mult=(1, 2), num_res_blocks=2, attn_resolutions=[], resolution=32,
                                      z_channels=3, in_channels=3, double_z=False, dropout=0.0), embed_dim=3)
        vae.load_state_dict(torch.load(ckpt)["state_dict"], strict=False)        # a Lightning VQModel checkpoint
        ldm = LatentDiffusion(unet, vae, latent_shape=(3, 16, 16), ...)
    """

    def __init__(self, ddconfig: dict, embed_dim: int, **ignored):
        super().__init__()
        cfg = dict(ddconfig)
        cfg.pop("in_channels", None); cfg.pop("double_z", None); cfg.pop("dropout", None)
        self.ddconfig, self.embed_dim = dict(ddconfig), embed_dim
        self._shapes = decoder_param_shapes(embed_dim=embed_dim, **cfg)
        for name, (shape, kind) in self._shapes.items():
            val = torch.zeros(shape) if kind == "zeros" else _init_param(shape, kind)
            _register(self, name, nn.Parameter(val, requires_grad=False))
        self._engines: Dict[tuple, VaeDecodeEngine] = {}
        self._versions = None

    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        self._engines.clear()
        return super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def engine(self, batch: int, h: int, w: int, device, lib=None) -> VaeDecodeEngine:
        v = self._weights_version()
        if v != self._versions:            # weights were replaced / moved / updated in place: re-pack
            self._engines.clear()
            self._versions = v
        key = (batch, h, w, str(device))
        if key not in self._engines:
            self._engines[key] = VaeDecodeEngine(dict(self.named_parameters()), batch, h, w, device, lib=lib)
        return self._engines[key]

    @torch.no_grad()
    def decode(self, quant: torch.Tensor) -> torch.Tensor:
        """autoencoder.py:113-116: fp32 NCHW latents -> fp32 NCHW images."""
        b, _, h, w = quant.shape
        eng = self.engine(b, h, w, quant.device)
        eng.z.copy_(quant)
        eng.run()
        return eng.out.clone()

    def encode(self, x):
        raise NotImplementedError("encoding (VQModel.encode: Encoder + quantiser, autoencoder.py:102-106) is not on the sampling "
                                  "path; encode with the reference's VQModel and pass the latents")

    def forward(self, quant):
        return self.decode(quant)
