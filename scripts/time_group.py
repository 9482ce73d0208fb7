"""Times a run of consecutive ops of the B = 1024 engine back to back (no L2 flush inside the group; one flush before it), to see
how much of the per-op cold time survives when the ops follow each other as in the sampling loop.
usage: python scripts/time_group.py <first-tag-substring> <count>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusion_models_b200 as ddm

B = int(os.environ.get("B", "1024"))
model = ddm.Unet(dim=64, dim_mults=(1, 2, 4, 8)).cuda().eval()
eng = model.engine(B, 32, 32, time_rows=1)
s = torch.cuda.current_stream().cuda_stream
for _, op in eng.time_ops:
    op(s)
eng.run_body(s)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tags = [t for t, _ in eng.ops]
i0 = next(i for i, t in enumerate(tags) if sys.argv[1] in t)
n = int(sys.argv[2])
ts = []
for i in range(12):
    flush.fill_(i & 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _, op in eng.ops[i0:i0 + n]:
        op(s)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(tags[i0:i0 + n], f"{sorted(ts[2:])[5]:.1f} us for the group")
