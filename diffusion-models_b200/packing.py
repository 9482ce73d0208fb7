"""Weight re-layout for the sm_100a kernels (host-side, pure tensor reshapes on the fp32 master weights).

The conv kernel consumes bf16 weight matrices [N_pad][K_pad] (K contiguous) whose K axis is ordered
(tap, source, channel), each (tap, source) segment zero-padded to a multiple of 64 channels so that one 64-wide
k-chunk never straddles two segments.  Everything below produces (matrix, taps) pairs in that convention.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

CHUNK = 64


def _ceil(v: int, m: int) -> int:
    return (v + m - 1) // m * m


@dataclass
class PackedConv:
    weight: torch.Tensor                 # bf16 [N_pad, K_pad]
    taps: List[Tuple[int, int, int]]     # (dy, dx, p) per tap
    seg_channels: Tuple[int, ...]        # channels per source as seen by the kernel (2C for the unshuffle view)
    n: int
    view: int = 0

    @property
    def n_pad(self) -> int:
        return self.weight.shape[0]

    @property
    def k_pad(self) -> int:
        return self.weight.shape[1]


def _assemble(per_tap: Sequence[Sequence[torch.Tensor]], n: int) -> torch.Tensor:
    """per_tap[t][s] is the fp32 [N, C_s] slice for tap t / source s."""
    cols = []
    for segs in per_tap:
        for w in segs:
            c = w.shape[1]
            pad = _ceil(c, CHUNK) - c
            cols.append(torch.nn.functional.pad(w, (0, pad)) if pad else w)
    m = torch.cat(cols, dim=1)
    n_pad = _ceil(n, 16)
    if n_pad != n:
        m = torch.nn.functional.pad(m, (0, 0, 0, n_pad - n))
    return m.to(torch.bfloat16).contiguous()


def pack_conv(weight: torch.Tensor, split: Optional[Sequence[int]] = None,
              in_scale: Optional[torch.Tensor] = None) -> PackedConv:
    """k x k stride-1 'same' convolution (nn.Conv2d weight [N, C, k, k]); `split` = channels per concatenated source;
    `in_scale` [C] is folded into the input-channel axis (RMSNorm gain of a preceding pre-norm)."""
    w = weight.detach().float()
    n, c, kh, kw = w.shape
    if in_scale is not None:
        w = w * in_scale.detach().float().reshape(1, c, 1, 1)
    split = tuple(split) if split else (c,)
    assert sum(split) == c
    taps, per_tap = [], []
    for ky in range(kh):
        for kx in range(kw):
            taps.append((ky - kh // 2, kx - kw // 2, 0))
            segs, o = [], 0
            for s in split:
                segs.append(w[:, o:o + s, ky, kx])
                o += s
            per_tap.append(segs)
    return PackedConv(_assemble(per_tap, n), taps, split, n)


def append_shortcut(pk: PackedConv, res_weight: torch.Tensor, split: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Weights of a conv with the block's 1x1 shortcut fused (ddm_conv_args.rsrc0, ResnetBlock.res_conv dd:134): the packed
    shortcut matrix [N_pad, sum of 64-padded source segments] appended to the conv's packed matrix along K."""
    r = pack_conv(res_weight, split)
    assert r.n_pad == pk.n_pad and len(r.taps) == 1
    return torch.cat([pk.weight, r.weight], dim=1).contiguous()


def pack_linear(weight: torch.Tensor) -> PackedConv:
    """nn.Linear weight [N, K] as a 1x1 convolution over a token matrix."""
    w = weight.detach().float()
    return PackedConv(_assemble([[w]], w.shape[0]), [(0, 0, 0)], (w.shape[1],), w.shape[0])


def pack_downsample(weight: torch.Tensor) -> PackedConv:
    """Rearrange('b c (h p1) (w p2) -> b (c p1 p2) h w') + 1x1 conv (denoising_diffusion.py:54-58) as a two-tap conv
    over the pixel-unshuffle *view* of the source: tap p1 reads source row 2y+p1, whose 2C contiguous values are
    ordered (p2, c).  Reference input channel index = c*4 + p1*2 + p2."""
    w = weight.detach().float()
    n, c4 = w.shape[:2]
    c = c4 // 4
    w = w.reshape(n, c, 2, 2)                      # [n, c, p1, p2]
    per_tap = [[w[:, :, p1, :].permute(0, 2, 1).reshape(n, 2 * c)] for p1 in range(2)]
    return PackedConv(_assemble(per_tap, n), [(0, 0, 0), (0, 0, 1)], (2 * c,), n, view=1)


def pack_upsample(weight: torch.Tensor) -> List[Tuple[PackedConv, int, int]]:
    """nn.Upsample(2, 'nearest') + 3x3 conv (denoising_diffusion.py:48-52) as four sub-pixel phases.

    Output pixel (2i+ph, 2j+pw) only sees source rows {i-1, i} (ph=0) or {i, i+1} (ph=1) -- likewise for columns --
    so each phase is a 2x2-tap convolution on the *un-upsampled* source with pre-summed weights, stored at stride 2.
    This removes the upsampled intermediate and 5/9 of the multiply-adds; zero padding is unchanged because the
    upsampled border maps to the source border.
    """
    w = weight.detach().float()
    n = w.shape[0]
    rows = {0: [(-1, (0,)), (0, (1, 2))], 1: [(0, (0, 1)), (1, (2,))]}   # phase -> [(source offset, kernel rows summed)]
    out = []
    for ph in range(2):
        for pw in range(2):
            taps, per_tap = [], []
            for dy, kys in rows[ph]:
                for dx, kxs in rows[pw]:
                    acc = torch.zeros_like(w[:, :, 0, 0])
                    for ky in kys:
                        for kx in kxs:
                            acc = acc + w[:, :, ky, kx]
                    taps.append((dy, dx, 0))
                    per_tap.append([acc])
            out.append((PackedConv(_assemble(per_tap, n), taps, (w.shape[1],), n), ph, pw))
    return out


def pack_stem(weight: torch.Tensor) -> torch.Tensor:
    """init_conv weight [Cout, Cin, k, k] -> fp32 [(ky, kx, ci)][Cout] for the direct stem kernel."""
    w = weight.detach().float()
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 1, 0).reshape(kh * kw * ci, co).contiguous()


def norm_gain(g: torch.Tensor) -> torch.Tensor:
    """RMSNorm gain with the sqrt(C) factor folded in (denoising_diffusion.py:60-67)."""
    g = g.detach().float().reshape(-1)
    return (g * (g.numel() ** 0.5)).contiguous()


def linattn_k_shift(to_qkv_weight: torch.Tensor, norm_g: torch.Tensor, mem_kv: torch.Tensor, heads: int, dim_head: int) -> torch.Tensor:
    """Per-channel shift for the fused linear attention's softmax over the tokens (denoising_diffusion.py:185).

    k[c][token] = w_c . x_hat with x_hat = x / ||x|| a unit vector (the block's RMSNorm, its gain g * sqrt(C) folded into
    w_c), so |k| <= ||w_c||_2; the learned memory keys (mem_kv[0], :181) are constants.  softmax is invariant to the
    shift, so exp(k - bound) / sum replaces the running maximum.  Returns fp32 [heads * dim_head]."""
    hid = heads * dim_head
    w = pack_conv(to_qkv_weight, in_scale=norm_gain(norm_g)).weight.float()[hid:2 * hid]      # the bf16 values the kernel multiplies
    bound = w.norm(dim=1) * 1.01 + 1e-3
    mem = mem_kv.detach().float()
    if mem.shape[-1] > 0:
        bound = torch.maximum(bound, mem[0].reshape(hid, -1).max(dim=1).values.to(bound.device))
    return bound.contiguous()
