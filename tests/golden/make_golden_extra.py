"""Extra golden fixtures from the UNMODIFIED reference (build container only): `DenoisingDiffusion.interpolate`
(dd:785-803).  Same conventions as make_golden.py (stubs, synthetic weights, every randn draw captured)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, load_synth, rnd, save, CaptureRandn   # noqa: E402


@torch.inference_mode()
def main():
    torch.set_num_threads(os.cpu_count())
    dd, _, _ = import_reference()
    base = dd.Unet(dim=64, dim_mults=(1, 2, 4, 8))
    load_synth(base, 0)
    d = dd.DenoisingDiffusion(base, image_size=32, timesteps=6, beta_schedule="cosine")
    x1, x2 = rnd((2, 3, 32, 32), 81).clamp(-1, 1), rnd((2, 3, 32, 32), 82).clamp(-1, 1)
    torch.manual_seed(1234)
    with CaptureRandn() as cap:
        y = d.interpolate(x1, x2, t=4, lam=0.3)
    # draws: q_sample noise for x1, for x2 (dd:793), then one randn_like per p_sample with t > 0 (dd:643)
    save("interpolate_T6", x1=x1, x2=x2, y=y, q_noise=torch.stack(cap.draws[:2]), noises=torch.stack(cap.draws[2:]))
    print("draws", len(cap.draws))

    # LearnedGaussianDiffusion (learned_gaussian_diffusion.py:60-111): ancestral sampling with the interpolated
    # log-variance.  (Its model_predictions / DDIM path references undefined names upstream and cannot run.)
    import denoising_diffusion.learned_gaussian_diffusion as lg
    lv = dd.Unet(dim=64, dim_mults=(1, 2, 4, 8), learned_variance=True)
    shapes = load_synth(lv, 12)
    ld = lg.LearnedGaussianDiffusion(lv, image_size=32, timesteps=6, beta_schedule="cosine")
    torch.manual_seed(1234)
    with CaptureRandn() as cap:
        y = ld.p_sample_loop((2, 3, 32, 32))
    save("learned_var_T6", y=y, x_T=cap.draws[0], noises=torch.stack(cap.draws[1:]))
    x, t = rnd((2, 3, 32, 32), 83), torch.tensor([5, 1])
    save("unet_learned_var_32", x=x, t=t, y=lv(x, t))
    import json
    with open(os.path.join(HERE, "manifest_extra.json"), "w") as f:
        json.dump({"unet_learned_var_32": {k: list(v) for k, v in shapes.items()}}, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
