"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs only in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

It imports the reference package (with two stub modules for `ema_pytorch` and
`accelerate`, which the reference imports at module top but uses only in its
Trainer), loads deterministic synthetic weights (`oracle.weights`), runs the
reference's own `Unet.forward`, `ddim_sample` and `p_sample_loop`, and writes
inputs + outputs as small .npz files.  Nothing here is imported by the product
or needed on the GPU box; the fixtures are what travels.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/denoising-diffusion-pytorch"
sys.path.insert(0, ROOT)


def import_reference():
    for name, attr in (("ema_pytorch", "EMA"), ("accelerate", "Accelerator")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            setattr(m, attr, type(attr, (), {}))
            sys.modules[name] = m
    sys.path.insert(0, REF)
    import denoising_diffusion.denoising_diffusion as dd
    import denoising_diffusion.denoising_diffusion_image_conditional as ic
    import denoising_diffusion.denoising_diffusion_text_conditional as tc
    return dd, ic, tc


def load_synth(model, seed):
    from oracle.weights import synth_state_dict
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synth_state_dict(shapes, seed), strict=True)
    model.eval()
    return shapes


def rnd(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


def save(name, **arrs):
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


class CaptureRandn:
    """Record every torch.randn / randn_like draw the reference sampler makes (dd:651,643,676,697)."""

    def __enter__(self):
        self.draws = []
        self._randn, self._randn_like = torch.randn, torch.randn_like

        def randn(*a, **k):
            t = self._randn(*a, **k)
            self.draws.append(t.clone())
            return t

        def randn_like(x, **k):
            t = self._randn_like(x, **k)
            self.draws.append(t.clone())
            return t

        torch.randn, torch.randn_like = randn, randn_like
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like


@torch.inference_mode()
def main():
    torch.set_num_threads(os.cpu_count())
    dd, ic, tc = import_reference()
    manifest = {}

    # ---- U-Net forward cases -------------------------------------------------
    def unet_case(name, model, x, t, seed=0, **fw):
        manifest[name] = {k: list(s) for k, s in load_synth(model, seed).items()}
        y = model(x, t, **fw)
        save(name, x=x, t=t, y=y, **{k: v for k, v in fw.items() if torch.is_tensor(v)})

    base = dd.Unet(dim=64, dim_mults=(1, 2, 4, 8))
    unet_case("unet_base_32", base, rnd((2, 3, 32, 32), 1), torch.tensor([999, 17]))
    unet_case("unet_base_64", base, rnd((1, 3, 64, 64), 2), torch.tensor([500]))
    unet_case("unet_small_16", dd.Unet(dim=32, dim_mults=(1, 2), attn_heads=2, attn_dim_head=16),
              rnd((2, 3, 16, 16), 3), torch.tensor([3, 640]), seed=5)
    unet_case("unet_selfcond_32", dd.Unet(dim=64, dim_mults=(1, 2, 4, 8), self_condition=True),
              rnd((1, 3, 32, 32), 4), torch.tensor([250]), seed=6, x_self_cond=rnd((1, 3, 32, 32), 40))
    unet_case("unet_imgcond_32", ic.Unet(dim=64, dim_mults=(1, 2, 4, 8), channels=4, cond_channels=4),
              rnd((1, 4, 32, 32), 5), torch.tensor([777]), seed=7, cond=rnd((1, 4, 32, 32), 50))
    unet_case("unet_text_xattn_32", tc.Unet(dim=64, channels=4, text_condition=True, use_cross_attn=True),
              rnd((1, 4, 32, 32), 6), torch.tensor([123]), seed=8, text_emb=rnd((1, 77, 512), 60))
    unet_case("unet_text_concat_32", tc.Unet(dim=64, channels=4, text_condition=True, use_cross_attn=False),
              rnd((2, 4, 32, 32), 7), torch.tensor([5, 900]), seed=9, text_emb=rnd((2, 512), 70))
    unet_case("unet_full_attn_all_16", dd.Unet(dim=32, dim_mults=(1, 2), full_attn=(True, True)),
              rnd((1, 3, 16, 16), 8), torch.tensor([42]), seed=10)

    # ---- schedules ---------------------------------------------------------
    names = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
             "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
             "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]
    sched = {}
    for kind, T in (("linear", 1000), ("cosine", 1000), ("sigmoid", 1000), ("cosine", 6)):
        d = dd.DenoisingDiffusion(base, image_size=32, timesteps=T, beta_schedule=kind)
        for n in names:
            sched[f"{kind}_{T}_{n}"] = getattr(d, n)
    save("schedules", **sched)
    pairs = {f"T{T}_S{S}": np.asarray(
        list(zip(*(lambda ts: (ts[:-1], ts[1:]))(list(reversed(torch.linspace(-1, T - 1, steps=S + 1).int().tolist()))))))
        for T, S in ((1000, 100), (1000, 50), (1000, 30), (1000, 250), (8, 3))}
    save("ddim_pairs", **pairs)

    # ---- samplers (reference loops, every randn draw captured) ------------------
    load_synth(base, 0)

    def sampler_case(name, diff, call):
        torch.manual_seed(1234)
        with CaptureRandn() as cap:
            y = call(diff)
        save(name, y=y, x_T=cap.draws[0], noises=torch.stack(cap.draws[1:]) if len(cap.draws) > 1 else np.zeros(0))

    sampler_case("ddim_eta0_S5", dd.DenoisingDiffusion(base, image_size=32, sampling_timesteps=5),
                 lambda d: d.ddim_sample((2, 3, 32, 32)))
    sampler_case("ddim_eta1_S4", dd.DenoisingDiffusion(base, image_size=32, sampling_timesteps=4, ddim_sampling_eta=1.0),
                 lambda d: d.ddim_sample((2, 3, 32, 32), return_all_timesteps=True))
    sampler_case("ddpm_T6", dd.DenoisingDiffusion(base, image_size=32, timesteps=6, beta_schedule="cosine"),
                 lambda d: d.p_sample_loop((2, 3, 32, 32)))
    sampler_case("ddim_predv_S3", dd.DenoisingDiffusion(base, image_size=32, sampling_timesteps=3, objective="pred_v",
                                                       beta_schedule="cosine"),
                 lambda d: d.ddim_sample((1, 3, 32, 32)))
    sampler_case("sample_dispatch_S3", dd.DenoisingDiffusion(base, image_size=32, sampling_timesteps=3),
                 lambda d: d.sample(batch_size=2))

    # image-conditional DDIM, called the way SURVEY 8c prescribes (the sample() wrapper is broken upstream)
    icm = ic.Unet(dim=64, dim_mults=(1, 2, 4, 8), channels=4, cond_channels=4)
    load_synth(icm, 7)
    cond = rnd((1, 4, 32, 32), 51)
    icd = ic.ImageConditionalDenoisingDiffusion(icm, image_size=32, auto_normalize=False, sampling_timesteps=3,
                                                condition_data_folder=None)
    torch.manual_seed(1234)
    with CaptureRandn() as cap:
        y = icd.ddim_sample((1, 4, 32, 32), sampling_timesteps=3, cond=cond)
    save("ddim_imgcond_S3", y=y, x_T=cap.draws[0], cond=cond)

    # text-conditional (cross-attn) DDIM with the condition fetch monkey-patched to a synthetic tensor
    tcm = tc.Unet(dim=64, channels=4, text_condition=True, use_cross_attn=True)
    load_synth(tcm, 8)
    text = rnd((1, 77, 512), 61)
    tcd = tc.TextConditionalDenoisingDiffusion(model=tcm, embedding_file=__file__, image_size=32,
                                               auto_normalize=False, sampling_timesteps=3)
    tcd.get_random_text_condition = lambda batch, device: (text, ["synthetic"] * batch)
    torch.manual_seed(1234)
    with CaptureRandn() as cap:
        y = tcd.ddim_sample((1, 4, 32, 32), sampling_timesteps=3)
    save("ddim_text_xattn_S3", y=y, x_T=cap.draws[0], text_emb=text)

    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    print("wrote manifest.json with", {k: len(v) for k, v in manifest.items()})


if __name__ == "__main__":
    main()
