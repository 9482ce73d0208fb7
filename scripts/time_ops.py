"""Per-op timing of the B = 1024 benchmark engine: every op of the plan is timed on its own (CUDA events, median of 9 after 2
warm-ups, a 256 MiB flush write before each launch).  usage: python scripts/time_ops.py [substring ...]  (env B, IMG)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusion_models_b200 as ddm

B, IMG = int(os.environ.get("B", "1024")), int(os.environ.get("IMG", "32"))
model = ddm.Unet(dim=64, dim_mults=(1, 2, 4, 8)).cuda().eval()
eng = model.engine(B, IMG, IMG, time_rows=1)
s = torch.cuda.current_stream().cuda_stream
for _, op in eng.time_ops:
    op(s)
eng.run_body(s)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
want = sys.argv[1:]
total = 0.0
for tag, op in eng.ops:
    if want and not any(w in tag for w in want):
        continue
    ts = []
    for i in range(11):
        flush.fill_(i & 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); op(s); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts[2:])[4]
    total += us
    print(f"{tag:34s} {us:8.1f} us", flush=True)
print(f"{'sum':34s} {total:8.1f} us")
