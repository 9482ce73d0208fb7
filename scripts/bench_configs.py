"""Secondary configurations of SURVEY.md section 8(d) (not bench lines): images/s and conv+linear TFLOP/s for
C3 (image-conditional latent 4x64x64 + cond, B=256), C4 (text cross-attention latent 4x64x64, B=128) and the C5
pixel sweep, each timed over a short DDIM run (S steps) on one GPU and reported per U-Net evaluation."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusion_models_b200 as ddm
from diffusion_models_b200.image_conditional import Unet as ImgUnet, ImageConditionalDenoisingDiffusion
from diffusion_models_b200.text_conditional import Unet as TextUnet, TextConditionalDenoisingDiffusion
from diffusion_models_b200.flops import unet_flops_per_image

S = int(os.environ.get("S", "10"))


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, B, size, model, ms):
    gf = unet_flops_per_image(model.spec, size, size) / 1e9
    evals = B * S / (ms / 1e3)
    print("CONFIG " + json.dumps({"config": name, "batch": B, "image": size, "ddim_steps_timed": S, "ms": round(ms, 2),
                                  "unet_evals_per_s": round(evals, 1), "ddim100_images_per_s": round(evals / 100, 1),
                                  "gflop_per_eval": round(gf, 3), "tflops": round(evals * gf / 1e3, 1)}), flush=True)


def base(B, size):
    m = ddm.Unet(dim=64, dim_mults=(1, 2, 4, 8)).cuda().eval()
    d = ddm.DenoisingDiffusion(m, image_size=size, sampling_timesteps=S).cuda()
    x = torch.randn(B, 3, size, size, device="cuda")
    report(f"C5 base {size}px", B, size, m, timed(lambda: d.ddim_sample((B, 3, size, size), noise=x)))


def c3(B=256):
    m = ImgUnet(dim=64, dim_mults=(1, 2, 4, 8), channels=4, cond_channels=4).cuda().eval()
    d = ImageConditionalDenoisingDiffusion(m, image_size=64, auto_normalize=False, sampling_timesteps=S).cuda()
    x, c = torch.randn(B, 4, 64, 64, device="cuda"), torch.randn(B, 4, 64, 64, device="cuda")
    report("C3 image-cond latent", B, 64, m, timed(lambda: d.ddim_sample((B, 4, 64, 64), sampling_timesteps=S, cond=c, noise=x)))


def c4(B=128):
    m = TextUnet(dim=64, channels=4, text_condition=True, use_cross_attn=True).cuda().eval()
    d = TextConditionalDenoisingDiffusion(model=m, image_size=64, auto_normalize=False, sampling_timesteps=S).cuda()
    x, t = torch.randn(B, 4, 64, 64, device="cuda"), torch.randn(B, 77, 512, device="cuda")
    report("C4 text-xattn latent", B, 64, m, timed(lambda: d.ddim_sample((B, 4, 64, 64), sampling_timesteps=S, text_emb=t, noise=x)))


if __name__ == "__main__":
    only = os.environ.get("ONLY")
    fns = {"32": lambda: base(1024, 32), "64": lambda: base(256, 64), "128": lambda: base(64, 128), "c3": c3, "c4": c4}
    for key, fn in fns.items():
        if only and key not in only.split(","):
            continue
        try:
            fn()
        except Exception as e:          # report and go on: these are secondary configurations
            print("CONFIG ERROR", type(e).__name__, str(e)[:300], flush=True)
        torch.cuda.empty_cache()
