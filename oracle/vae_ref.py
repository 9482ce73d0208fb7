"""fp32 functional restatement of the VAE decode that follows the latent sampling loop (oracle; tests only).

`LD` = /root/reference/latent-diffusion/ldm:  VQModel.decode (`LD/models/autoencoder.py:113-116`) = post_quant_conv (1x1)
-> Decoder (`LD/modules/diffusionmodules/model.py:479-585`), built from ResnetBlock (:99-158, temb is None in the
autoencoder), AttnBlock (:167-219), Upsample (:59-74), Normalize = GroupNorm(32, eps=1e-6) (:55-56) and swish (:50-52).
Driven by a reference-format state_dict (`decoder.*`, `post_quant_conv.*`)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def _norm(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)              # model.py:55-56


def _swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)                                                          # model.py:50-52


def resnet_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """model.py:115-138 with temb = None (the autoencoder's decoder has temb_ch = 0)."""
    h = F.conv2d(_swish(_norm(sd, p + ".norm1", x)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(_swish(_norm(sd, p + ".norm2", h)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    elif p + ".conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"], padding=1)
    return x + h


def attn_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """model.py:190-215 -- single-head attention over the h*w positions, scale c^-0.5, residual."""
    h_ = _norm(sd, p + ".norm", x)
    q = F.conv2d(h_, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv2d(h_, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv2d(h_, sd[p + ".v.weight"], sd[p + ".v.bias"])
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + F.conv2d(h_, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])


def decoder_forward(sd: SD, z: Tensor, p: str = "decoder") -> Tensor:
    """model.py:552-585 (give_pre_end = False, tanh_out = False)."""
    h = F.conv2d(z, sd[p + ".conv_in.weight"], sd[p + ".conv_in.bias"], padding=1)
    h = resnet_block(sd, p + ".mid.block_1", h)
    if p + ".mid.attn_1.q.weight" in sd:
        h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    levels = 0
    while f"{p}.up.{levels}.block.0.conv1.weight" in sd:
        levels += 1
    for lvl in reversed(range(levels)):
        j = 0
        while f"{p}.up.{lvl}.block.{j}.conv1.weight" in sd:
            h = resnet_block(sd, f"{p}.up.{lvl}.block.{j}", h)
            if f"{p}.up.{lvl}.attn.{j}.q.weight" in sd:
                h = attn_block(sd, f"{p}.up.{lvl}.attn.{j}", h)
            j += 1
        if f"{p}.up.{lvl}.upsample.conv.weight" in sd:                                  # model.py:70-73 nearest 2x + conv
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"{p}.up.{lvl}.upsample.conv.weight"], sd[f"{p}.up.{lvl}.upsample.conv.bias"], padding=1)
    h = _swish(_norm(sd, p + ".norm_out", h))
    return F.conv2d(h, sd[p + ".conv_out.weight"], sd[p + ".conv_out.bias"], padding=1)


def vae_decode(sd: SD, z: Tensor) -> Tensor:
    """autoencoder.py:113-116."""
    return decoder_forward(sd, F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]))
