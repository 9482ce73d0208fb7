"""Algorithmic work of one U-Net evaluation, counted the way SURVEY.md section 8(d) defines it: 2 x MAC of every
nn.Conv2d / nn.Linear the *reference* module executes (forward-hook count), attention einsums excluded.  Used by
bench.py for `roofline.achieved`; independent of how our kernels restructure the work (e.g. the 4-phase upsample
does 4/9 of the reference MACs -- it is still credited with the reference's count, no more)."""
from __future__ import annotations

from .arch import UnetSpec


def unet_flops_per_image(spec: UnetSpec, height: int, width: int, text_tokens: int = 0) -> float:
    total = 0.0
    conv = lambda co, ci, k, h, w: 2.0 * co * ci * k * k * h * w
    lin = lambda co, ci, rows=1: 2.0 * co * ci * rows
    td = spec.time_dim
    h, w = height, width
    total += conv(spec.init_dim, spec.input_channels, spec.stem_kernel, h, w)
    total += lin(td, spec.fourier_dim) + lin(td, td)

    def resblock(rb, h, w):
        f = lin(2 * rb.c_out, td) + conv(rb.c_out, rb.c_in, 3, h, w) + conv(rb.c_out, rb.c_out, 3, h, w)
        if rb.c_in != rb.c_out:
            f += conv(rb.c_out, rb.c_in, 1, h, w)
        return f

    def attn(at, h, w):
        hid = at.heads * at.dim_head
        return conv(3 * hid, at.dim, 1, h, w) + conv(at.dim, hid, 1, h, w)

    for st in spec.downs:
        total += resblock(st.block1, h, w) + resblock(st.block2, h, w) + attn(st.attn, h, w)
        if st.resample_kind == "down":
            h, w = h // 2, w // 2
            total += conv(st.c_res_out, 4 * st.c_res_in, 1, h, w)
        else:
            total += conv(st.c_res_out, st.c_res_in, 3, h, w)
    total += resblock(spec.mid1, h, w) + attn(spec.mid_attn, h, w) + resblock(spec.mid2, h, w)
    if spec.text_mode == "xattn":
        inner, mid, m = spec.xattn_heads * spec.xattn_dim_head, spec.mid1.c_out, max(text_tokens, 1)
        total += 3 * (lin(inner, mid, h * w) + 2 * lin(inner, spec.text_emb_dim, m) + lin(mid, inner, h * w))
    elif spec.text_mode == "concat":
        total += lin(td, spec.text_emb_dim) + lin(td, td) + lin(td, 2 * td)
    for st in spec.ups:
        total += resblock(st.block1, h, w) + resblock(st.block2, h, w) + attn(st.attn, h, w)
        if st.resample_kind == "up":
            h, w = 2 * h, 2 * w
        total += conv(st.c_res_out, st.c_res_in, 3, h, w)
    total += resblock(spec.final_block, h, w) + conv(spec.out_dim, spec.init_dim, 1, h, w)
    return total
