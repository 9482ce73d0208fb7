"""One eager U-Net evaluation (uniform timestep, sampling configuration) for ncu: prints the op tags in launch order."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusion_models_b200 as ddm
from oracle import synth_state_dict

B = int(os.environ.get("B", "1024"))
S = int(os.environ.get("IMG", "32"))
REPS = int(os.environ.get("REPS", "3"))
model = ddm.Unet(dim=64, dim_mults=(1, 2, 4, 8))
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 0))
model = model.cuda().eval()
eng = model.engine(B, S, S, time_rows=1)
eng.x.normal_()
eng.time.fill_(500.0)
eng.run_time_path()
for _ in range(REPS):
    eng.run_body()
torch.cuda.synchronize()
tags = [t for t, _ in eng.ops]
out = os.path.join(ROOT, "gpurun_out", "op_tags.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump({"B": B, "img": S, "time_ops": [t for t, _ in eng.time_ops], "body_ops": tags, "meta": eng.op_meta}, open(out, "w"))
print(len(eng.time_ops), "time launches +", len(tags), "body launches per forward")
