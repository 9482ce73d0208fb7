"""Device-side timeline of the fused linear-attention kernel (CTA 0: control warp and epilogue warp 0).
Run with DDM_LAF_TRACE=1.  usage: laf_trace.py [B] [n]"""
import ctypes as C, os, sys, statistics
sys.path.insert(0, ".")
os.environ.setdefault("DDM_LAF_TRACE", "1")
import torch
from diffusion_models_b200 import _lib
from diffusion_models_b200._lib import LinAttnBlockArgs
from diffusion_models_b200.packing import linattn_k_shift, norm_gain, pack_conv

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
lib = _lib.init(0)
g = torch.Generator().manual_seed(0)
C_, heads, d = 64, 4, 32
hid = heads * d
x = torch.randn((B, n, C_), generator=g).to("cuda", torch.bfloat16)
w_qkv = torch.randn((3 * hid, C_, 1, 1), generator=g) * 0.125
g_in = torch.ones((1, C_, 1, 1))
w_out = torch.randn((C_, hid, 1, 1), generator=g) * 0.09
mem = torch.randn((2, heads, d, 4), generator=g)
keep = [pack_conv(w_qkv, in_scale=norm_gain(g_in)).weight.cuda(), pack_conv(w_out).weight.cuda(), torch.zeros(C_, device="cuda"),
        norm_gain(torch.ones(1, C_, 1, 1)).cuda(), mem.cuda(), linattn_k_shift(w_qkv, g_in, mem, heads, d).cuda()]
out = torch.zeros_like(x)
a = LinAttnBlockArgs()
a.x, a.out, a.B, a.n, a.C = x.data_ptr(), out.data_ptr(), B, n, C_
a.w_qkv, a.w_out, a.bias_out, a.g_out, a.mem_kv, a.k_shift = (t.data_ptr() for t in keep)
a.heads, a.dim_head, a.n_mem = heads, d, 4
s = torch.cuda.current_stream().cuda_stream
buf = (C.c_longlong * (3 * 8192))()
for _ in range(2):
    _lib.check(lib.ddm_linear_attention_block(C.byref(a), s)); torch.cuda.synchronize()
    cnt = lib.ddm_debug_linattn_trace(buf, 8192)
ev = sorted((buf[3 * i + 2], buf[3 * i], buf[3 * i + 1] >> 32, buf[3 * i + 1] & 0xFFFFFFFF) for i in range(cnt))
t0 = ev[0][0]
names = {9: "kv ready to issue", 10: "kv issued", 11: "edone ok", 12: "ctx issued", 20: "q loop top", 21: "qdone ok", 22: "y issued", 23: "ydone ok",
         30: "cdone ok", 31: "m issued", 32: "mtdone ok", 1: "x ok", 2: "rn bar", 3: "acc ok", 4: "pv ok", 5: "epi1 math done", 6: "arrived",
         40: "x ok(2)", 41: "rn bar(2)", 42: "acc ok(2)", 43: "q epi done", 44: "q arrived", 45: "y ok", 46: "y red bar", 47: "y arrived",
         50: "ctx complete", 51: "cdone arrived", 52: "mdone ok", 53: "mt arrived"}
lo, hi = int(os.environ.get("FROM", "0")), int(os.environ.get("TO", "60000"))
print(f"TRACE B={B} n={n}: {cnt} events, span {ev[-1][0] - t0} cycles")
for t, role, e, idx in ev:
    if lo <= t - t0 <= hi:
        print(f"TRACE {t - t0:8d} {'control ' if role == 0 else 'epilogue'} {names.get(e, e):18s} #{idx}")
