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


if __name__ == "__main__":
    main()
