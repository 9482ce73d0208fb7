"""Time selected conv layers of the plan in isolation (CUDA events), optionally under DDM_CONV_DEBUG bisection flags."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusion_models_b200 as ddm

B = int(os.environ.get("B", "1024"))
tags = os.environ.get("TAGS", "downs.0.0.block1,downs.0.0.block2,ups.3.0.block1,downs.0.2.to_qkv,downs.0.2.to_out,ups.2.0.block1,ups.1.0.block1,mid_block1.block1.gemm").split(",")
model = ddm.Unet(dim=64, dim_mults=(1, 2, 4, 8)).cuda().eval()
eng = model.engine(B, 32, 32, time_rows=1)
ops = dict(eng.ops)
s = torch.cuda.current_stream().cuda_stream
for i in range(2):
    eng.run_body(s)
torch.cuda.synchronize()
for tag in tags:
    op = ops[tag]
    for _ in range(3):
        op(s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        op(s)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    m = eng.op_meta.get(tag)
    tf = 2.0 * m["M"] * m["N"] * m["K"] / us / 1e6 if m else 0
    print(f"DBG={os.environ.get('DDM_CONV_DEBUG','0')} {tag:28s} {us:8.1f} us  {tf:7.1f} TF", flush=True)
