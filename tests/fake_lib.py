"""A stand-in for libddm_b200.so that executes each C-ABI call with the torch statements in tests/kernel_ref.py.

TEST INFRASTRUCTURE ONLY (CPU, `-m "not gpu"`): it lets `UnetEngine` build and run its real plan -- real packed
weights, real ddm_conv_args structs, real raw pointers -- without a GPU, so the host logic is checked against the
oracle before any GPU time is spent.  The product never imports this module.
"""
import ctypes as C

import torch

import kernel_ref as R


class FakeLib:
    def __init__(self):
        self.engine = None
        self.calls = 0
        self.num_sms = 148

    def attach(self, engine):
        self.engine = engine

    # ---- pointer -> tensor view -------------------------------------------------------------------
    def _tensors(self):
        e = self.engine
        for t in e._keep:
            if torch.is_tensor(t):
                yield t
        for v in vars(e).values():
            if torch.is_tensor(v):
                yield v
            elif isinstance(v, dict):
                for u in v.values():
                    if torch.is_tensor(u):
                        yield u
                    elif isinstance(u, (list, tuple)):
                        for w in u:
                            if torch.is_tensor(w):
                                yield w
        for t in getattr(self, "extra", []):
            yield t

    def view(self, ptr, shape, dtype):
        if ptr is None or ptr == 0:
            return None
        n = 1
        for s in shape:
            n *= s
        for t in self._tensors():
            base = t.data_ptr()
            size = t.numel() * t.element_size()
            if base <= ptr < base + size and t.dtype == dtype:
                off = (ptr - base) // t.element_size()
                flat = t.reshape(-1)
                assert off + n <= flat.numel(), "view overruns its tensor"
                return flat[off:off + n].view(*shape)
        raise KeyError(f"pointer {ptr:#x} ({dtype}) not found in engine tensors")

    def strided_rows(self, ptr, rows, ld, cols, dtype):
        """rows x cols window with row stride ld starting at ptr."""
        if ptr is None or ptr == 0:
            return None
        for t in self._tensors():
            base = t.data_ptr()
            size = t.numel() * t.element_size()
            if base <= ptr < base + size and t.dtype == dtype:
                off = (ptr - base) // t.element_size()
                flat = t.reshape(-1)
                return torch.as_strided(flat, (rows, cols), (ld, 1), off)
        raise KeyError(f"pointer {ptr:#x} not found")

    # ---- entry points -----------------------------------------------------------------------------------
    def ddm_conv2d(self, ref, stream):
        a = ref._obj
        self.calls += 1
        bf, f32 = torch.bfloat16, torch.float32
        B, H, W = a.B, a.H, a.W
        hs, ws = (2 * H, 2 * W) if a.view == 1 else (H, W)
        srcs = [self.strided_rows(a.src0, B * hs * ws, a.ld0, a.C0, bf).float().reshape(B, hs, ws, a.C0)]
        if a.src1:
            srcs.append(self.strided_rows(a.src1, B * hs * ws, a.ld1, a.C1, bf).float().reshape(B, hs, ws, a.C1))
        weight = self.view(a.weight, (a.N_pad, a.K_pad), bf).float()
        taps = [(a.tap_dy[i], a.tap_dx[i], a.tap_p[i]) for i in range(a.ntaps)]
        shortcut = None
        if a.rsrc0:        # fused 1x1 shortcut: its weights are the last rC0 + rC1 columns of K
            rs = [self.strided_rows(a.rsrc0, B * H * W, a.rld0, a.rC0, bf).float()]
            if a.rsrc1:
                rs.append(self.strided_rows(a.rsrc1, B * H * W, a.rld1, a.rC1, bf).float())
            kr = a.rC0 + a.rC1
            shortcut = torch.cat(rs, dim=1) @ weight[:a.N, a.K_pad - kr:].t()
            if a.rbias:
                shortcut = shortcut + self.view(a.rbias, (a.N,), f32)
            shortcut = shortcut.reshape(B, H, W, a.N)
            weight = weight[:, :a.K_pad - kr].contiguous()
        N = a.N
        row_scale = self.view(a.row_scale, (B * H * W,), f32)
        bias = self.view(a.bias, (N,), f32)
        g = self.view(a.norm_g, (N,), f32)
        ss = None
        if a.scale_shift:
            rows = B if a.ss_stride else 1
            ss = self.strided_rows(a.scale_shift, rows, max(a.ss_stride, 2 * N), 2 * N, f32)
        if a.head_out:             # fused head: the tile's fp32 values go through final_conv instead of being stored
            res = shortcut
            if a.residual:
                res = self.strided_rows(a.residual, B * H * W, a.ld_res, a.ld_res, bf).float().reshape(B, H, W, a.ld_res)
            v = torch.zeros((B, H, W, N), dtype=f32)
            R.conv_ref(srcs, weight, N, (B, H, W), taps, view=a.view, row_scale=row_scale, bias=bias, norm_g=g, scale_shift=ss, act=a.act,
                       residual=res, out=v, round_out=False)
            hw = self.view(a.head_w, (a.head_n, N), f32)
            y = v.reshape(-1, N) @ hw.t() + self.view(a.head_b, (a.head_n,), f32)
            self.view(a.head_out, (B, a.head_n, H, W), f32).copy_(y.reshape(B, H, W, a.head_n).permute(0, 3, 1, 2))
            return 0
        if a.out_f32_nchw:
            out = self.view(a.out, (B, N, a.OH, a.OW), f32)
            res = None
        else:
            out = self.strided_rows(a.out, B * a.OH * a.OW, a.ld_out, a.ld_out, bf)
            outf = out.float().reshape(B, a.OH, a.OW, a.ld_out)
            res = None
            if a.residual:
                res = self.strided_rows(a.residual, B * a.OH * a.OW, a.ld_res, a.ld_res, bf).float().reshape(B, a.OH, a.OW, a.ld_res)
            elif shortcut is not None:
                res = shortcut
        rn = self.view(a.rnorm_out, (B * a.OH * a.OW,), f32)
        if a.ksplit > 1:           # split-K: raw sums; the emulation puts the whole sum into range 0 and zeros elsewhere
            raw = torch.zeros((B, H, W, N), dtype=f32)
            R.conv_ref(srcs, weight, N, (B, H, W), taps, view=a.view, out=raw, round_out=False)
            part = self.view(a.partial_out, (a.ksplit, B * H * W, N), f32)
            part.zero_()
            part[0].copy_(raw.reshape(B * H * W, N))
            return 0
        if a.out_f32_nchw:
            R.conv_ref(srcs, weight, N, (B, H, W), taps, view=a.view, row_scale=row_scale, bias=bias, norm_g=g,
                       scale_shift=ss, act=a.act, out=out, out_map=(a.sy, a.sx, a.oy, a.ox), out_f32_nchw=True)
        else:
            R.conv_ref(srcs, weight, N, (B, H, W), taps, view=a.view, row_scale=row_scale, bias=bias, norm_g=g,
                       scale_shift=ss, act=a.act, residual=res, out=outf, out_map=(a.sy, a.sx, a.oy, a.ox), rnorm_out=rn)
            out.copy_(outf.reshape(out.shape).to(bf))
        return 0

    def ddm_conv2d_shortcut_supported(self, N, C_in, rC0, rC1, H, W):
        if N not in (64, 128) or C_in != N or rC0 < 64 or rC0 % 64 or rC1 % 64 or rC0 + rC1 > 256:
            return 0
        bw = 32 if W >= 32 else 1 << max(W - 1, 0).bit_length()
        if bw * H < 128:
            return 0
        return 1 if (W % bw == 0 and H % (128 // bw) == 0) else 0

    def ddm_stem_conv(self, in0, c0, in1, c1, in2, c2, w, b, out, B, H, W, Cout, ks, stream):
        self.calls += 1
        f32 = torch.float32
        ins = [self.view(p, (B, c, H, W), f32) for p, c in ((in0, c0), (in1, c1), (in2, c2)) if c]
        cin = c0 + c1 + c2
        y = R.stem_ref(ins, self.view(w, (ks * ks * cin, Cout), f32), self.view(b, (Cout,), f32), ks, Cout)
        self.view(out, (B, H, W, Cout), torch.bfloat16).copy_(y.to(torch.bfloat16))
        return 0

    def ddm_head_conv1x1(self, x, w, b, out, B, HW, C_, N, stream):
        self.calls += 1
        xv = self.view(x, (B, HW, C_), torch.bfloat16).float()
        y = xv @ self.view(w, (N, C_), torch.float32).t() + self.view(b, (N,), torch.float32)
        self.view(out, (B, N, HW), torch.float32).copy_(y.permute(0, 2, 1))
        return 0

    def ddm_sinusoidal_embedding(self, t, out, rows, dim, theta, stream):
        self.calls += 1
        self.view(out, (rows, dim), torch.float32).copy_(R.sinusoidal_ref(self.view(t, (rows,), torch.float32), dim, theta))
        return 0

    def ddm_small_linear(self, x, ldx, W, b, y, ldy, rows, N, K, act_in, act_out, stream):
        self.calls += 1
        f32 = torch.float32
        xv = self.strided_rows(x, rows, ldx, K, f32)
        out = R.small_linear_ref(xv, self.view(W, (N, K), f32), self.view(b, (N,), f32), act_in, act_out)
        self.strided_rows(y, rows, ldy, N, f32).copy_(out)
        return 0

    def ddm_row_rnorm(self, x, ld, rn, rows, C_, stream):
        self.calls += 1
        self.view(rn, (rows,), torch.float32).copy_(R.row_rnorm_ref(self.strided_rows(x, rows, ld, C_, torch.bfloat16)))
        return 0

    def ddm_rmsnorm_act(self, x, g, ss, ss_stride, rows_per_batch, act, res, out, rows, C_, stream):
        self.calls += 1
        bf, f32 = torch.bfloat16, torch.float32
        ssv = None
        if ss:
            nb = (rows + rows_per_batch - 1) // rows_per_batch if ss_stride else 1
            ssv = self.strided_rows(ss, nb, max(ss_stride, 2 * C_), 2 * C_, f32)
        r = self.view(res, (rows, C_), bf)
        y = R.rmsnorm_act_ref(self.view(x, (rows, C_), bf).float(), self.view(g, (C_,), f32), ssv, rows_per_batch, act,
                              r.float() if r is not None else None)
        self.view(out, (rows, C_), bf).copy_(y.to(bf))
        return 0

    def ddm_conv2d_head_supported(self, N, head_n, H, W):
        return 1 if (N in (64, 128) and 1 <= head_n <= 4) else 0

    def ddm_conv2d_row_norm_supported(self, N):
        return 1 if (N <= 256 or (N <= 512 and N % 128 == 0)) else 0

    def ddm_conv2d_suggest_ksplit(self, rows, N_pad, K_pad):
        m_tiles, n_tiles, stages = (rows + 127) // 128, (N_pad + 255) // 256, K_pad // 64
        tiles = m_tiles * n_tiles
        if tiles * 3 > self.num_sms or stages < 8:
            return 1
        ks = min(self.num_sms // tiles, stages // 4, 16)
        while ks > 1 and (ks - 1) * ((stages + ks - 1) // ks) >= stages:
            ks -= 1
        return max(ks, 1)

    def ddm_rmsnorm_act_split(self, part, ksplit, bias, g, ss, ss_stride, rows_per_batch, act, res, out, rows, C_, stream):
        self.calls += 1
        bf, f32 = torch.bfloat16, torch.float32
        x = self.view(part, (ksplit, rows, C_), f32).sum(0)
        if bias:
            x = x + self.view(bias, (C_,), f32)
        ssv = None
        if ss:
            nb = (rows + rows_per_batch - 1) // rows_per_batch if ss_stride else 1
            ssv = self.strided_rows(ss, nb, max(ss_stride, 2 * C_), 2 * C_, f32)
        r = self.view(res, (rows, C_), bf)
        y = R.rmsnorm_act_ref(x, self.view(g, (C_,), f32), ssv, rows_per_batch, act, r.float() if r is not None else None)
        self.view(out, (rows, C_), bf).copy_(y.to(bf))
        return 0

    def ddm_groupnorm_act(self, x, gamma, beta, out, B, HW, C_, groups, eps, act, stream):
        self.calls += 1
        bf, f32 = torch.bfloat16, torch.float32
        xv = self.view(x, (B, HW, C_), bf).float().permute(0, 2, 1)                       # [B, C, HW]
        y = torch.nn.functional.group_norm(xv, groups, self.view(gamma, (C_,), f32), self.view(beta, (C_,), f32), eps=eps)
        if act == 1:
            y = y * torch.sigmoid(y)
        self.view(out, (B, HW, C_), bf).copy_(y.permute(0, 2, 1).to(bf))
        return 0

    def ddm_linear_attention(self, qkv, mem_kv, out, B, n, heads, d, n_mem, stream):
        self.calls += 1
        bf = torch.bfloat16
        y = R.linear_attention_ref(self.view(qkv, (B, n, 3 * heads * d), bf).float(),
                                   self.view(mem_kv, (2, heads, d, n_mem), torch.float32), heads, d)
        self.view(out, (B, n, heads * d), bf).copy_(y.to(bf))
        return 0

    def ddm_linear_attention_bounded(self, qkv, mem_kv, k_shift, out, B, n, heads, d, n_mem, stream):
        return self.ddm_linear_attention(qkv, mem_kv, out, B, n, heads, d, n_mem, stream)      # (the shift does not change the softmax)

    def ddm_linear_attention_block_supported(self, C_, n, heads, d, n_mem):
        return 1 if (C_ in (64, 128) and heads == 4 and d == 32 and n >= 128 and n % 128 == 0 and 0 <= n_mem <= 4) else 0

    def ddm_linear_attention_block(self, ref, stream):
        a = ref._obj
        self.calls += 1
        bf, f32 = torch.bfloat16, torch.float32
        hid = a.heads * a.dim_head
        x = self.view(a.x, (a.B, a.n, a.C), bf).float()
        y = R.linattn_block_ref(x, self.view(a.w_qkv, (3 * hid, a.C), bf).float(), self.view(a.w_out, (a.C, hid), bf).float(),
                                self.view(a.bias_out, (a.C,), f32), self.view(a.g_out, (a.C,), f32),
                                self.view(a.mem_kv, (2, a.heads, a.dim_head, a.n_mem), f32), a.heads, a.dim_head)
        self.view(a.out, (a.B, a.n, a.C), bf).copy_(y.to(bf))
        return 0

    def ddm_attention(self, q, ldq, k, ldk, v, ldv, mem_k, mem_v, n_mem, out, B, nq, nk, heads, d, stream):
        self.calls += 1
        bf, f32 = torch.bfloat16, torch.float32
        hd = heads * d
        qv = self.strided_rows(q, B * nq, ldq, hd, bf).float().reshape(B, nq, hd)
        kv = self.strided_rows(k, B * nk, ldk, hd, bf).float().reshape(B, nk, hd)
        vv = self.strided_rows(v, B * nk, ldv, hd, bf).float().reshape(B, nk, hd)
        mk = self.view(mem_k, (heads, n_mem, d), f32) if n_mem else None
        mv = self.view(mem_v, (heads, n_mem, d), f32) if n_mem else None
        self.view(out, (B, nq, hd), bf).copy_(R.attention_ref(qv, kv, vv, mk, mv, heads, d).to(bf))
        return 0
