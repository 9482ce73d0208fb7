"""fp32 functional restatement of the reference U-Net forward (oracle; tests only).

Every function cites the reference lines it restates.  `R` below abbreviates
/root/reference/denoising-diffusion-pytorch/denoising_diffusion/ :
  dd = R/denoising_diffusion.py, ic = R/denoising_diffusion_image_conditional.py,
  tc = R/denoising_diffusion_text_conditional.py, at = R/attend.py.

The network is driven purely by a reference-format ``state_dict`` (the
interchange format) plus a small `UnetConfig`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class UnetConfig:
    """Hyper-parameters that are not recoverable from tensor shapes alone."""
    heads: Tuple[int, ...] = (4, 4, 4, 4)          # dd:289-295 attn_heads per stage
    dim_head: Tuple[int, ...] = (32, 32, 32, 32)   # dd:289-295 attn_dim_head per stage
    theta: float = 10000.0                         # dd:276 sinusoidal_pos_emb_theta
    self_condition: bool = False                   # dd:259-260
    xattn_heads: int = 4                           # tc:120-125 (hard-coded heads=4)
    n_stages: int = 4
    full_attn: Tuple[bool, ...] = field(default_factory=tuple)


def infer_config(sd: SD, *, heads=4, dim_head=32, theta=10000.0, self_condition=False) -> UnetConfig:
    """Recover stage count / attention kinds from the key set (dd:299-341)."""
    n = 0
    while f"downs.{n}.0.block1.proj.weight" in sd:
        n += 1
    full = tuple(f"downs.{i}.2.to_out.weight" in sd for i in range(n))
    as_t = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v,) * n
    return UnetConfig(heads=as_t(heads), dim_head=as_t(dim_head), theta=theta,
                      self_condition=self_condition, n_stages=n, full_attn=full)


# ----------------------------------------------------------------------------
# leaf ops
# ----------------------------------------------------------------------------

def rms_norm(x: Tensor, g: Tensor) -> Tensor:
    """dd:60-67 -- F.normalize over channels (eps 1e-12 clamps the L2 norm) * g * sqrt(C)."""
    c = x.shape[1]
    nrm = x.pow(2).sum(dim=1, keepdim=True).sqrt().clamp_min(1e-12)
    return x / nrm * g * (c ** 0.5)


def rms_norm_1d(x: Tensor, g: Tensor) -> Tensor:
    """tc:27-36 -- same over the last dim of (B, n, C); g is (1, C)."""
    c = x.shape[-1]
    nrm = x.pow(2).sum(dim=-1, keepdim=True).sqrt().clamp_min(1e-12)
    return x / nrm * g * (c ** 0.5)


def sinusoidal_emb(t: Tensor, dim: int, theta: float) -> Tensor:
    """dd:77-84 -- cat(sin(t f), cos(t f)), f_j = exp(-j ln(theta)/(half-1))."""
    half = dim // 2
    step = math.log(theta) / (half - 1)
    f = torch.exp(torch.arange(half, device=t.device) * -step)
    a = t[:, None] * f[None, :]
    return torch.cat((a.sin(), a.cos()), dim=-1)


def time_mlp(sd: SD, t: Tensor, theta: float) -> Tensor:
    """dd:280-285 -- sinusoid -> Linear -> exact-erf GELU -> Linear."""
    fourier_dim = sd["time_mlp.1.weight"].shape[1]
    e = sinusoidal_emb(t, fourier_dim, theta)
    e = F.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    e = F.gelu(e)
    return F.linear(e, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def block(sd: SD, p: str, x: Tensor, scale_shift=None) -> Tensor:
    """dd:113-122 -- conv3x3 -> RMSNorm -> x*(scale+1)+shift -> SiLU (dropout is identity in eval)."""
    x = F.conv2d(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"], padding=1)
    x = rms_norm(x, sd[p + ".norm.g"])
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def resnet_block(sd: SD, p: str, x: Tensor, t_emb: Tensor) -> Tensor:
    """dd:136-148 -- mlp(SiLU->Linear) chunk -> block1(ss) -> block2 -> + res_conv(x)."""
    ss = F.linear(F.silu(t_emb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])
    ss = ss[:, :, None, None]
    scale, shift = ss.chunk(2, dim=1)
    h = block(sd, p + ".block1", x, (scale, shift))
    h = block(sd, p + ".block2", h)
    if p + ".res_conv.weight" in sd:
        res = F.conv2d(x, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])
    else:
        res = x
    return h + res


def linear_attention(sd: SD, p: str, x: Tensor, heads: int) -> Tensor:
    """dd:173-193 -- RMSNorm, qkv 1x1, mem-kv first, q softmax over d, k softmax over n, to_out conv + RMSNorm."""
    b, c, hh, ww = x.shape
    xn = rms_norm(x, sd[p + ".norm.g"])
    qkv = F.conv2d(xn, sd[p + ".to_qkv.weight"])
    q, k, v = (z.reshape(b, heads, -1, hh * ww) for z in qkv.chunk(3, dim=1))
    mem = sd[p + ".mem_kv"]                                    # (2, h, d, n_mem)
    mk = mem[0][None].expand(b, -1, -1, -1)
    mv = mem[1][None].expand(b, -1, -1, -1)
    k = torch.cat((mk, k), dim=-1)
    v = torch.cat((mv, v), dim=-1)
    d = q.shape[2]
    q = q.softmax(dim=-2) * (d ** -0.5)
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q)
    out = out.reshape(b, heads * d, hh, ww)
    out = F.conv2d(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return rms_norm(out, sd[p + ".to_out.1.g"])


def full_attention(sd: SD, p: str, x: Tensor, heads: int) -> Tensor:
    """dd:215-229 + at:109-124 -- RMSNorm, qkv, mem-kv rows first, softmax(q k^T d^-0.5) v, to_out conv."""
    b, c, hh, ww = x.shape
    xn = rms_norm(x, sd[p + ".norm.g"])
    qkv = F.conv2d(xn, sd[p + ".to_qkv.weight"])
    q, k, v = (z.reshape(b, heads, -1, hh * ww).transpose(-1, -2) for z in qkv.chunk(3, dim=1))
    mem = sd[p + ".mem_kv"]                                    # (2, h, n_mem, d)
    k = torch.cat((mem[0][None].expand(b, -1, -1, -1), k), dim=-2)
    v = torch.cat((mem[1][None].expand(b, -1, -1, -1), v), dim=-2)
    d = q.shape[-1]
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * (d ** -0.5)
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
    out = out.transpose(-1, -2).reshape(b, heads * d, hh, ww)
    return F.conv2d(out, sd[p + ".to_out.weight"], sd[p + ".to_out.bias"])


def cross_attention(sd: SD, p: str, x: Tensor, ctx: Tensor, heads: int) -> Tensor:
    """tc:54-78 -- q from x (B,n,C), k/v from text (B,m,Ct); softmax over m; Linear + RMSNorm1D."""
    if ctx.ndim == 2:
        ctx = ctx[:, None]
    b, n, _ = x.shape
    m = ctx.shape[1]
    q = F.linear(x, sd[p + ".to_q.weight"]).reshape(b, n, heads, -1).transpose(1, 2)
    k = F.linear(ctx, sd[p + ".to_k.weight"]).reshape(b, m, heads, -1).transpose(1, 2)
    v = F.linear(ctx, sd[p + ".to_v.weight"]).reshape(b, m, heads, -1).transpose(1, 2)
    d = q.shape[-1]
    att = (torch.einsum("bhnd,bhmd->bhnm", q, k) * (d ** -0.5)).softmax(dim=-1)
    out = torch.einsum("bhnm,bhmd->bhnd", att, v).transpose(1, 2).reshape(b, n, heads * d)
    out = F.linear(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return rms_norm_1d(out, sd[p + ".to_out.1.g"])


def downsample(sd: SD, p: str, x: Tensor) -> Tensor:
    """dd:54-58 -- 'b c (h p1) (w p2) -> b (c p1 p2) h w' then 1x1 conv; dd:319 plain 3x3 on the last stage."""
    if p + ".1.weight" in sd:
        b, c, hh, ww = x.shape
        x = x.reshape(b, c, hh // 2, 2, ww // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(b, c * 4, hh // 2, ww // 2)
        return F.conv2d(x, sd[p + ".1.weight"], sd[p + ".1.bias"])
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)


def upsample(sd: SD, p: str, x: Tensor) -> Tensor:
    """dd:48-52 -- nearest 2x then 3x3 conv; dd:336 plain 3x3 on the last stage."""
    if p + ".1.weight" in sd:
        x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        return F.conv2d(x, sd[p + ".1.weight"], sd[p + ".1.bias"], padding=1)
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)


# ----------------------------------------------------------------------------
# whole network
# ----------------------------------------------------------------------------

def unet_forward(sd: SD, x: Tensor, time: Tensor, cfg: Optional[UnetConfig] = None, *,
                 x_self_cond: Optional[Tensor] = None, cond: Optional[Tensor] = None,
                 text_emb: Optional[Tensor] = None, taps: Optional[dict] = None) -> Tensor:
    """dd:349-390 (+ ic:51-55 image condition, tc:131-214 text condition).

    `taps`, if given, is filled with named intermediate activations so that
    per-layer parity tests can localise an error.
    """
    cfg = cfg or infer_config(sd)
    n = cfg.n_stages
    assert all(d % (2 ** (n - 1)) == 0 for d in x.shape[-2:]), "dd:350 spatial dims must divide the downsample factor"
    rec = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)

    if cond is not None:                                    # ic:52-54 (x first, cond second)
        assert cond.shape[0] == x.shape[0]
        x = torch.cat((x, cond), dim=1)
    if cfg.self_condition:                                  # dd:352-354 (self-cond first)
        sc = x_self_cond if x_self_cond is not None else torch.zeros_like(x)
        x = torch.cat((sc, x), dim=1)

    x = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)   # dd:356
    r = x                                                                        # dd:357
    rec("init_conv", x)
    t = time_mlp(sd, time, cfg.theta)                                            # dd:359

    use_xattn = text_emb is not None and "cross_attn.to_q.weight" in sd
    if text_emb is not None and "text_proj.0.weight" in sd:                      # tc:146-152
        te = text_emb
        if te.dim() == 3 and te.size(1) == 1:
            te = te.squeeze(1)
        te = te.to(t.dtype)
        f = F.linear(te, sd["text_proj.0.weight"], sd["text_proj.0.bias"])
        f = F.linear(F.gelu(f), sd["text_proj.2.weight"], sd["text_proj.2.bias"])
        t = F.linear(torch.cat((t, f), dim=1), sd["text_concat_proj.weight"], sd["text_concat_proj.bias"])
    rec("t_emb", t)

    def attn(p, z, i):
        if cfg.full_attn[i]:
            return full_attention(sd, p, z, cfg.heads[i])
        return linear_attention(sd, p, z, cfg.heads[i])

    def xattn(p, z):                                                             # tc:173-177
        b, c, hs, ws = z.shape
        flat = z.reshape(b, c, hs * ws).permute(0, 2, 1)
        flat = cross_attention(sd, p, flat, text_emb, cfg.xattn_heads)
        return flat.permute(0, 2, 1).reshape(b, c, hs, ws)

    skips = []
    for i in range(n):                                                           # dd:363-371
        x = resnet_block(sd, f"downs.{i}.0", x, t); skips.append(x); rec(f"downs.{i}.0", x)
        x = resnet_block(sd, f"downs.{i}.1", x, t); rec(f"downs.{i}.1", x)
        x = attn(f"downs.{i}.2", x, i) + x; skips.append(x); rec(f"downs.{i}.2", x)
        x = downsample(sd, f"downs.{i}.3", x); rec(f"downs.{i}.3", x)

    if use_xattn:
        x = xattn("cross_attn_down", x); rec("cross_attn_down", x)
    x = resnet_block(sd, "mid_block1", x, t); rec("mid_block1", x)               # dd:373
    if use_xattn:
        x = xattn("cross_attn", x); rec("cross_attn", x)                         # tc:183-188
    x = full_attention(sd, "mid_attn", x, cfg.heads[-1]) + x; rec("mid_attn", x) # dd:374
    x = resnet_block(sd, "mid_block2", x, t); rec("mid_block2", x)               # dd:375
    if use_xattn:
        x = xattn("cross_attn_up", x); rec("cross_attn_up", x)                   # tc:194-198

    for j in range(n):                                                           # dd:377-385
        i = n - 1 - j                                                            # reversed stage kinds (dd:327)
        x = torch.cat((x, skips.pop()), dim=1)
        x = resnet_block(sd, f"ups.{j}.0", x, t); rec(f"ups.{j}.0", x)
        x = torch.cat((x, skips.pop()), dim=1)
        x = resnet_block(sd, f"ups.{j}.1", x, t); rec(f"ups.{j}.1", x)
        x = attn(f"ups.{j}.2", x, i) + x; rec(f"ups.{j}.2", x)
        x = upsample(sd, f"ups.{j}.3", x); rec(f"ups.{j}.3", x)

    x = torch.cat((x, r), dim=1)                                                 # dd:387
    x = resnet_block(sd, "final_res_block", x, t); rec("final_res_block", x)     # dd:389
    return F.conv2d(x, sd["final_conv.weight"], sd["final_conv.bias"])           # dd:390
