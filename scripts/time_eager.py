"""Host cost of the eager path: wall-clock of Unet.forward (B = 16, 32x32) per call, CPU launch loop + GPU, averaged."""
import sys, time, torch
sys.path.insert(0, ".")
import diffusion_models_b200 as ddm
m = ddm.Unet(dim=64, dim_mults=(1, 2, 4, 8)).cuda().eval()
x, t = torch.randn(16, 3, 32, 32, device="cuda"), torch.full((16,), 500, device="cuda")
for _ in range(5):
    m(x, t)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    m(x, t)
torch.cuda.synchronize()
print(f"eager Unet.forward B=16: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per call")
