"""Static description of the reference U-Net: which parameters exist (names/shapes in the reference's own
state_dict format) and how the layers are wired.  One description drives both the nn.Module parameter registration
(`unet.Unet`) and the kernel plan (`engine.UnetEngine`).

Reference: denoising-diffusion-pytorch/denoising_diffusion/denoising_diffusion.py:233-390 (Unet),
denoising_diffusion_image_conditional.py:31-55, denoising_diffusion_text_conditional.py:86-214.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


def _tuple(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


@dataclass
class ResBlockSpec:
    name: str            # state_dict prefix, e.g. "downs.0.0"
    c_in: int            # channels of the (possibly concatenated) input
    c_out: int
    split: Optional[Tuple[int, int]] = None   # (current, skip) channel split when the input is a torch.cat


@dataclass
class AttnSpec:
    name: str
    dim: int
    heads: int
    dim_head: int
    full: bool           # softmax attention (dd:196-229) vs linear attention (dd:150-193)
    n_mem: int = 4


@dataclass
class StageSpec:
    block1: ResBlockSpec
    block2: ResBlockSpec
    attn: AttnSpec
    resample: str        # state_dict prefix of the stage-end conv
    resample_kind: str   # "down" (unshuffle + 1x1), "up" (nearest 2x + 3x3) or "conv" (plain 3x3)
    c_res_in: int
    c_res_out: int


@dataclass
class UnetSpec:
    dim: int
    init_dim: int
    out_dim: int
    channels: int
    input_channels: int      # channels seen by init_conv (self-cond and image-cond included)
    cond_channels: int
    self_condition: bool
    time_dim: int
    fourier_dim: int
    theta: float
    stem_kernel: int
    downs: List[StageSpec]
    mid1: ResBlockSpec
    mid_attn: AttnSpec
    mid2: ResBlockSpec
    ups: List[StageSpec]
    final_block: ResBlockSpec
    text_mode: Optional[str]     # None | "concat" | "xattn"
    text_emb_dim: int
    xattn_heads: int
    xattn_dim_head: int
    params: Dict[str, Tuple[Tuple[int, ...], str]] = field(default_factory=dict)   # name -> (shape, init kind)

    @property
    def n_stages(self) -> int:
        return len(self.downs)

    @property
    def downsample_factor(self) -> int:
        return 2 ** (len(self.downs) - 1)

    def res_blocks(self) -> List[ResBlockSpec]:
        """All ResnetBlocks in execution order (their scale/shift projections share one table)."""
        out = []
        for s in self.downs:
            out += [s.block1, s.block2]
        out += [self.mid1, self.mid2]
        for s in self.ups:
            out += [s.block1, s.block2]
        out.append(self.final_block)
        return out


def build_spec(dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=3, self_condition=False,
               learned_variance=False, sinusoidal_pos_emb_theta=10000, attn_dim_head=32, attn_heads=4, full_attn=None,
               cond_channels=0, text_mode=None, text_emb_dim=512, xattn_dim_head=32) -> UnetSpec:
    n = len(dim_mults)
    init_dim = init_dim if init_dim is not None else dim
    dims = [init_dim] + [dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    time_dim = dim * 4
    # the conv kernel reads channels in 16-byte units and tiles C_out up to 1024 (csrc/conv_tc.cuh kMaxNPad): say so at
    # construction instead of failing inside the first engine run
    bad = [c for c in dims if c % 8 != 0 or c > 1024]
    if bad:
        raise ValueError(f"Unet widths {dims}: every stage width must be a multiple of 8 and at most 1024 on the B200 path")
    if not full_attn:                                        # dd:289-290 full attention only in the innermost stage
        full_attn = (False,) * (n - 1) + (True,)
    full_attn, heads, dim_head = _tuple(full_attn, n), _tuple(attn_heads, n), _tuple(attn_dim_head, n)
    assert len(full_attn) == n
    input_channels = channels * (2 if self_condition else 1) + cond_channels
    P: Dict[str, Tuple[Tuple[int, ...], str]] = {}

    def conv(name, co, ci, k, bias=True):
        P[name + ".weight"] = ((co, ci, k, k), "conv")
        if bias:
            P[name + ".bias"] = ((co,), "bias:%d" % (ci * k * k))

    def linear(name, co, ci, bias=True):
        P[name + ".weight"] = ((co, ci), "conv")
        if bias:
            P[name + ".bias"] = ((co,), "bias:%d" % ci)

    def resblock(name, ci, co, split=None):
        linear(name + ".mlp.1", co * 2, time_dim)
        conv(name + ".block1.proj", co, ci, 3)
        P[name + ".block1.norm.g"] = ((1, co, 1, 1), "ones")
        conv(name + ".block2.proj", co, co, 3)
        P[name + ".block2.norm.g"] = ((1, co, 1, 1), "ones")
        if ci != co:
            conv(name + ".res_conv", co, ci, 1)
        return ResBlockSpec(name, ci, co, split)

    def attn(name, c, h, d, full):
        hid = h * d
        P[name + ".norm.g"] = ((1, c, 1, 1), "ones")
        P[name + ".mem_kv"] = ((2, h, 4, d) if full else (2, h, d, 4), "randn")
        conv(name + ".to_qkv", hid * 3, c, 1, bias=False)
        if full:
            conv(name + ".to_out", c, hid, 1)
        else:
            conv(name + ".to_out.0", c, hid, 1)
            P[name + ".to_out.1.g"] = ((1, c, 1, 1), "ones")
        return AttnSpec(name, c, h, d, full)

    conv("init_conv", init_dim, input_channels, 7)
    linear("time_mlp.1", time_dim, dim)
    linear("time_mlp.3", time_dim, time_dim)

    downs = []
    for i, (ci, co) in enumerate(in_out):
        last = i >= n - 1
        b1 = resblock(f"downs.{i}.0", ci, ci)
        b2 = resblock(f"downs.{i}.1", ci, ci)
        at = attn(f"downs.{i}.2", ci, heads[i], dim_head[i], full_attn[i])
        if last:
            conv(f"downs.{i}.3", co, ci, 3)
            downs.append(StageSpec(b1, b2, at, f"downs.{i}.3", "conv", ci, co))
        else:
            conv(f"downs.{i}.3.1", co, ci * 4, 1)
            downs.append(StageSpec(b1, b2, at, f"downs.{i}.3.1", "down", ci, co))

    mid = dims[-1]
    mid1 = resblock("mid_block1", mid, mid)
    mid_attn = attn("mid_attn", mid, heads[-1], dim_head[-1], True)
    mid2 = resblock("mid_block2", mid, mid)

    ups = []
    for j in range(n):
        i = n - 1 - j
        ci, co = in_out[i]
        last = j == n - 1
        b1 = resblock(f"ups.{j}.0", co + ci, co, split=(co, ci))
        b2 = resblock(f"ups.{j}.1", co + ci, co, split=(co, ci))
        at = attn(f"ups.{j}.2", co, heads[i], dim_head[i], full_attn[i])
        if last:
            conv(f"ups.{j}.3", ci, co, 3)
            ups.append(StageSpec(b1, b2, at, f"ups.{j}.3", "conv", co, ci))
        else:
            conv(f"ups.{j}.3.1", ci, co, 3)
            ups.append(StageSpec(b1, b2, at, f"ups.{j}.3.1", "up", co, ci))

    default_out = channels * (2 if learned_variance else 1)
    out_dim = out_dim if out_dim is not None else default_out
    final_block = resblock("final_res_block", init_dim * 2, init_dim, split=(init_dim, init_dim))
    conv("final_conv", out_dim, init_dim, 1)

    if text_mode == "concat":                                 # tc:107-114
        linear("text_proj.0", time_dim, text_emb_dim)
        linear("text_proj.2", time_dim, time_dim)
        linear("text_concat_proj", time_dim, time_dim * 2)
    elif text_mode == "xattn":                                # tc:120-125 (heads hard-coded to 4)
        inner = 4 * xattn_dim_head
        for nm in ("cross_attn", "cross_attn_down", "cross_attn_up"):
            linear(nm + ".to_q", inner, mid, bias=False)
            linear(nm + ".to_k", inner, text_emb_dim, bias=False)
            linear(nm + ".to_v", inner, text_emb_dim, bias=False)
            linear(nm + ".to_out.0", mid, inner)
            P[nm + ".to_out.1.g"] = ((1, mid), "ones")

    return UnetSpec(dim=dim, init_dim=init_dim, out_dim=out_dim, channels=channels, input_channels=input_channels,
                    cond_channels=cond_channels, self_condition=self_condition, time_dim=time_dim, fourier_dim=dim,
                    theta=float(sinusoidal_pos_emb_theta), stem_kernel=7, downs=downs, mid1=mid1, mid_attn=mid_attn,
                    mid2=mid2, ups=ups, final_block=final_block, text_mode=text_mode, text_emb_dim=text_emb_dim,
                    xattn_heads=4, xattn_dim_head=xattn_dim_head, params=P)
