"""Times the fused LinearAttention block kernel (ddm_linear_attention_block) at the benchmark shapes: CUDA events,
L2 flushed between launches, median of 20.  Prints us per launch and achieved GB/s on the algorithmic bytes
(one read of x + one write of y)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import diffusion_models_b200 as ddm
from diffusion_models_b200 import _lib
from diffusion_models_b200._lib import LinAttnBlockArgs
from diffusion_models_b200.packing import linattn_k_shift, norm_gain, pack_conv


def main():
    lib = _lib.init(0)
    g = torch.Generator().manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shapes = ((1024, 1024, 64), (1024, 256, 64), (256, 4096, 64), (64, 16384, 64), (128, 1024, 64),
              (1024, 256, 128), (1024, 1024, 128), (128, 256, 128))
    if len(sys.argv) > 2:
        shapes = ((int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[4]) if len(sys.argv) > 4 else 64),)
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    for B, n, C_ in shapes:
        heads, d = 4, 32
        hid = heads * d
        x = torch.randn((B, n, C_), generator=g).to("cuda", torch.bfloat16)
        w_qkv = torch.randn((3 * hid, C_, 1, 1), generator=g) / C_ ** 0.5
        g_in = torch.ones((1, C_, 1, 1))
        w_out = torch.randn((C_, hid, 1, 1), generator=g) * 0.09
        mem = torch.randn((2, heads, d, 4), generator=g)
        keep = [pack_conv(w_qkv, in_scale=norm_gain(g_in)).weight.cuda(), pack_conv(w_out).weight.cuda(),
                torch.zeros(C_, device="cuda"), norm_gain(torch.ones(1, C_, 1, 1)).cuda(), mem.cuda(),
                linattn_k_shift(w_qkv, g_in, mem, heads, d).cuda()]
        out = torch.zeros_like(x)
        a = LinAttnBlockArgs()
        a.x, a.out, a.B, a.n, a.C = x.data_ptr(), out.data_ptr(), B, n, C_
        a.w_qkv, a.w_out, a.bias_out, a.g_out, a.mem_kv, a.k_shift = (t.data_ptr() for t in keep)
        a.heads, a.dim_head, a.n_mem = heads, d, 4
        s = torch.cuda.current_stream().cuda_stream
        ts = []
        for i in range(reps):
            flush.fill_(i & 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.ddm_linear_attention_block(C.byref(a), s))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = sorted(ts[min(5, reps - 1):])
        us = ts[len(ts) // 2]
        gb = 2 * x.numel() * 2 / us / 1e3
        print(f"fused linattn block B={B} n={n} C={C_}: {us:8.1f} us   {gb:7.1f} GB/s algorithmic (x read + y write)", flush=True)


if __name__ == "__main__":
    main()
