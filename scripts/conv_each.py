"""Run every conv op of the B=1024 engine one by one (sync after each) and report the first that fails."""
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
torch.cuda.synchronize()
for tag, op in eng.ops:
    op(s)
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print("EACH FAILED at", tag, str(e)[:80]); sys.exit(1)
    print("EACH ok", tag)
