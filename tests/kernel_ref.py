"""Plain-torch statement of what each C-ABI entry point of include/ddm_b200.h computes (test infrastructure).

Used two ways: (1) `-m gpu` tests compare every kernel against these functions on random inputs; (2) the CPU
plan-emulation test (tests/fake_lib.py) executes a whole `UnetEngine` plan with these functions and compares it with
the oracle, which validates weight packing, tap tables, sub-pixel phases, views and the layer wiring without a GPU.
All math is fp32; tensors that are bf16 on the device are rounded to bf16 where the kernels round.
"""
import math

import torch
import torch.nn.functional as F


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _shift(src, dy, dx, H, W):
    """src [B,H,W,C] -> tensor whose (y,x) holds src(y+dy, x+dx), zero outside (== TMA out-of-bounds fill)."""
    B, Hs, Ws, C = src.shape
    out = src.new_zeros((B, H, W, C))
    y0, y1 = max(0, -dy), min(H, Hs - dy)
    x0, x1 = max(0, -dx), min(W, Ws - dx)
    if y1 > y0 and x1 > x0:
        out[:, y0:y1, x0:x1] = src[:, y0 + dy:y1 + dy, x0 + dx:x1 + dx]
    return out


def conv_ref(srcs, weight, N, domain, taps, *, view=0, row_scale=None, bias=None, norm_g=None, scale_shift=None,
             act=0, residual=None, out=None, out_map=(1, 1, 0, 0), out_f32_nchw=False, rnorm_out=None,
             round_out=True):
    """Semantics of ddm_conv2d.  srcs: list of float [B,Hs,Ws,C]; weight: [N_pad,K_pad] (values already bf16-rounded);
    scale_shift: [Bt, >=2N] rows (Bt in {1,B}); out: preallocated output tensor that is updated in place."""
    B, H, W = domain
    cols = []
    for (dy, dx, p) in taps:
        for s in srcs:
            if view == 0:
                a = _shift(s, dy, dx, H, W)
            else:                                   # [B,2H,2W,C] read as rows 2y+p, channel axis (p2, c)
                Bs, H2, W2, C = s.shape
                a = s.reshape(Bs, H2 // 2, 2, W2 // 2, 2 * C)[:, :, p]
                a = _shift(a, dy, dx, H, W)
            c = a.shape[-1]
            pad = (-c) % 64
            cols.append(F.pad(a, (0, pad)) if pad else a)
    A = torch.cat(cols, dim=-1)
    assert A.shape[-1] == weight.shape[1], (A.shape, weight.shape)
    v = A.reshape(-1, A.shape[-1]) @ weight[:N].float().t()
    v = v.reshape(B, H, W, N)
    if row_scale is not None:
        v = v * row_scale.reshape(B, H, W, 1)
    if bias is not None:
        v = v + bias
    if norm_g is not None:
        nrm = v.pow(2).sum(-1, keepdim=True).sqrt().clamp_min(1e-12)
        v = v / nrm * norm_g
    if scale_shift is not None:
        ss = scale_shift[:, :2 * N] if scale_shift.shape[0] == B else scale_shift[:1, :2 * N].expand(B, -1)
        v = v * (ss[:, None, None, :N] + 1) + ss[:, None, None, N:]
    if act == 1:
        v = F.silu(v)
    sy, sx, oy, ox = out_map
    if out_f32_nchw:
        out[:, :, oy::sy, ox::sx] = v.permute(0, 3, 1, 2)
        return out
    if residual is not None:
        v = v + residual[:, oy::sy, ox::sx, :N]
    if round_out:
        v = bf16_round(v)
    out[:, oy::sy, ox::sx, :N] = v
    if rnorm_out is not None:
        rn = 1.0 / v.pow(2).sum(-1).sqrt().clamp_min(1e-12)
        rnorm_out.reshape(out.shape[0], out.shape[1], out.shape[2])[:, oy::sy, ox::sx] = rn
    return out


def stem_ref(ins, weight_packed, bias, ksize, Cout):
    """ddm_stem_conv: ins = list of fp32 NCHW tensors (concatenated on C); weight [(ky,kx,ci)][Cout]."""
    x = torch.cat(ins, dim=1)
    ci = x.shape[1]
    w = weight_packed.reshape(ksize, ksize, ci, Cout).permute(3, 2, 0, 1)
    y = F.conv2d(x, w, bias, padding=ksize // 2)
    return bf16_round(y.permute(0, 2, 3, 1).contiguous())


def sinusoidal_ref(t, dim, theta):
    half = dim // 2
    f = torch.exp(torch.arange(half, device=t.device, dtype=torch.float32) * -(math.log(theta) / (half - 1)))
    a = t[:, None] * f[None]
    return torch.cat((a.sin(), a.cos()), dim=-1)


def _act(v, code):
    if code == 1:
        return F.silu(v)
    if code == 2:
        return F.gelu(v)
    return v


def small_linear_ref(x, W, b, act_in, act_out):
    y = _act(x, act_in) @ W.t()
    if b is not None:
        y = y + b
    return _act(y, act_out)


def row_rnorm_ref(x):
    return 1.0 / x.float().pow(2).sum(-1).sqrt().clamp_min(1e-12)


def rmsnorm_act_ref(x, g, ss, rows_per_batch, act, residual):
    """x [rows, C] float; ss [Bt, >=2C] or None."""
    rows, C = x.shape
    v = x
    if g is not None:
        v = v / v.pow(2).sum(-1, keepdim=True).sqrt().clamp_min(1e-12) * g
    if ss is not None:
        idx = torch.arange(rows, device=x.device) // rows_per_batch
        s = ss[idx] if ss.shape[0] > 1 else ss.expand(rows, -1)
        v = v * (s[:, :C] + 1) + s[:, C:2 * C]
    if act == 1:
        v = F.silu(v)
    if residual is not None:
        v = v + residual
    return bf16_round(v)


def linear_attention_ref(qkv, mem_kv, heads, d):
    """qkv [B,n,3*heads*d] float; mem_kv [2,heads,d,n_mem] -> [B,n,heads*d] (dd:178-192)."""
    B, n, _ = qkv.shape
    q, k, v = (z.reshape(B, n, heads, d).permute(0, 2, 3, 1) for z in qkv.chunk(3, dim=-1))   # b h d n
    k = torch.cat((mem_kv[0][None].expand(B, -1, -1, -1), k), dim=-1)
    v = torch.cat((mem_kv[1][None].expand(B, -1, -1, -1), v), dim=-1)
    q = q.softmax(dim=-2) * d ** -0.5
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q)
    return bf16_round(out.permute(0, 3, 1, 2).reshape(B, n, heads * d))


def linattn_block_ref(x, w_qkv, w_out, bias_out, g_out, mem_kv, heads, d):
    """ddm_linear_attention_block: x [B,n,C] float (bf16 values); w_qkv [3*heads*d, C] (pre-norm gain folded in),
    w_out [C, heads*d], g_out = g*sqrt(C); returns bf16-rounded RMSNorm(W_out attn + b) * g + x   (dd:173-193, :368)."""
    B, n, C = x.shape
    xn = x / x.pow(2).sum(-1, keepdim=True).sqrt().clamp_min(1e-12)
    qkv = xn @ w_qkv.t()
    q, k, v = (z.reshape(B, n, heads, d).permute(0, 2, 3, 1) for z in qkv.chunk(3, dim=-1))   # b h d n
    k = torch.cat((mem_kv[0][None].expand(B, -1, -1, -1), k), dim=-1)
    v = torch.cat((mem_kv[1][None].expand(B, -1, -1, -1), v), dim=-1)
    q = q.softmax(dim=-2) * d ** -0.5
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    o = torch.einsum("bhde,bhdn->bhen", ctx, q).permute(0, 3, 1, 2).reshape(B, n, heads * d)
    y = o @ w_out.t() + bias_out
    y = y / y.pow(2).sum(-1, keepdim=True).sqrt().clamp_min(1e-12) * g_out
    return bf16_round(y + x)


def attention_ref(q, k, v, mem_k, mem_v, heads, d):
    """q [B,nq,heads*d], k/v [B,nk,heads*d] float; mem_k/mem_v [heads,n_mem,d] or None."""
    B, nq, _ = q.shape
    nk = k.shape[1]
    qh = q.reshape(B, nq, heads, d).transpose(1, 2)
    kh = k.reshape(B, nk, heads, d).transpose(1, 2)
    vh = v.reshape(B, nk, heads, d).transpose(1, 2)
    if mem_k is not None:
        kh = torch.cat((mem_k[None].expand(B, -1, -1, -1), kh), dim=-2)
        vh = torch.cat((mem_v[None].expand(B, -1, -1, -1), vh), dim=-2)
    att = (qh @ kh.transpose(-1, -2) * d ** -0.5).softmax(dim=-1)
    return bf16_round((att @ vh).transpose(1, 2).reshape(B, nq, heads * d))


def sampler_step_ref(kind, x, mo, noise, coef, objective):
    """One row of the coefficient table applied like ddm_sampler_step (unfused fp32 ops in the reference's order).
    `coef` must be a tensor: the reference divides by a (B,1,1,1) *tensor* (true division, dd:576-580), whereas torch
    turns division by a Python scalar into a multiplication by its reciprocal."""
    ra, rm1, k2, k3, k4, k5, sac, s1m = (coef[i] for i in range(8))
    rax = ra * x
    if objective == 0:
        x0 = (rax - rm1 * mo).clamp(-1, 1)
        eps = (rax - x0) / rm1 if kind == 0 else mo
    else:
        x0 = mo.clamp(-1, 1) if objective == 1 else (sac * x - s1m * mo).clamp(-1, 1)
        eps = (rax - x0) / rm1
    z = noise if noise is not None else torch.zeros_like(x)
    if kind == 0:
        if k5 != 0:
            return x0, x0
        xn = x0 * k2 + k3 * eps
        if k4 != 0:
            xn = xn + k4 * z
        return xn, x0
    xn = k2 * x0 + k3 * x
    if k4 != 0:
        xn = xn + k4 * z
    return xn, x0
