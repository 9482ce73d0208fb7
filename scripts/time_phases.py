"""Where does one ddim_sample() call spend its time?  (diagnostic)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusion_models_b200 as ddm
from oracle import synth_state_dict

B = int(os.environ.get("B", "1024"))
model = ddm.Unet(dim=64, dim_mults=(1, 2, 4, 8))
model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 0))
model = model.cuda().eval()
diff = ddm.DenoisingDiffusion(model, image_size=32, sampling_timesteps=100).cuda()
x = torch.randn((B, 3, 32, 32), device="cuda")
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    y = diff.ddim_sample((B, 3, 32, 32), noise=x)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"call {i}: {1e3*(t1-t0):.1f} ms", flush=True)
# replay-only timing: drive the engine directly
eng = model.engine(B, 32, 32, time_rows=1)
s = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        eng.run_body(s)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"eager 20 forwards: {1e3*(t1-t0)/20:.2f} ms per forward", flush=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    eng.run_body(torch.cuda.current_stream().cuda_stream)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50):
        g.replay()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"graph 50 forwards: {1e3*(t1-t0)/50:.2f} ms per forward", flush=True)
