"""Round-2 golden fixtures from the UNMODIFIED reference (build container only; same conventions as make_golden.py):

  * U-Net forwards at the BASELINE shapes that the 32-px fixtures did not cover: image-conditional
    latent 4+4 x 64x64 (C3), text cross-attention latent 4 x 64x64 with 77 tokens (C4), unconditional 3 x 128x128 (C5);
  * `DenoisingDiffusion.ddim_sample_guided` (dd:710-777) with a guide / mask, eta = 0.5, every randn draw captured
    (its inline matplotlib display calls are stubbed out: matplotlib is not installed and is no part of the arithmetic).
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, load_synth, rnd, save, CaptureRandn   # noqa: E402


def stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    for name in ("figure", "subplot", "title", "imshow", "axis", "show"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt


@torch.inference_mode()
def main():
    torch.set_num_threads(os.cpu_count())
    dd, ic, tc = import_reference()

    def unet_case(name, model, x, t, seed, **fw):
        load_synth(model, seed)
        save(name, x=x, t=t, y=model(x, t, **fw), **{k: v for k, v in fw.items() if torch.is_tensor(v)})

    # same constructors / weight seeds as the 32-px fixtures, so manifest.json already pins their state_dict layout
    unet_case("unet_imgcond_64", ic.Unet(dim=64, dim_mults=(1, 2, 4, 8), channels=4, cond_channels=4),
              rnd((1, 4, 64, 64), 205), torch.tensor([640]), 7, cond=rnd((1, 4, 64, 64), 250))
    unet_case("unet_text_xattn_64", tc.Unet(dim=64, channels=4, text_condition=True, use_cross_attn=True),
              rnd((1, 4, 64, 64), 206), torch.tensor([321]), 8, text_emb=rnd((1, 77, 512), 260))
    base = dd.Unet(dim=64, dim_mults=(1, 2, 4, 8))
    unet_case("unet_base_128", base, rnd((1, 3, 128, 128), 207), torch.tensor([77]), 0)

    stub_matplotlib()
    load_synth(base, 0)
    d = dd.DenoisingDiffusion(base, image_size=32, sampling_timesteps=4, ddim_sampling_eta=0.5)
    guide = rnd((2, 3, 32, 32), 270).clamp(-1, 1)
    mask = (rnd((1, 1, 32, 32), 271) > 0).float().expand(2, 3, 32, 32).contiguous()
    torch.manual_seed(1234)
    with CaptureRandn() as cap:
        y = d.ddim_sample_guided((2, 3, 32, 32), guide=guide, mask=mask)
    # draws: x_T (dd:721), then per non-final step the DDIM noise (dd:741) and q_sample's noise for the guide (dd:748)
    steps = (len(cap.draws) - 1) // 2
    save("ddim_guided_S4", y=y, x_T=cap.draws[0], guide=guide, mask=mask,
         noises=torch.stack(cap.draws[1::2][:steps]), guide_noises=torch.stack(cap.draws[2::2][:steps]))
    print("guided draws", len(cap.draws))
    torch.manual_seed(1234)
    with CaptureRandn() as cap:
        y = d.ddim_sample_guided((2, 3, 32, 32), clip_denoised=False)
    save("ddim_guided_noguide_S4", y=y, x_T=cap.draws[0], noises=torch.stack(cap.draws[1:]))


# ---------------------------------------------------------------------------------------------------------------------
# VAE decode: the reference's Decoder + post_quant_conv, imported unmodified
# (ldm/modules/diffusionmodules/model.py needs only torch / numpy / einops; VQModel itself needs pytorch_lightning + taming,
# which are absent -- its decode() is exactly `decoder(post_quant_conv(quant))`, autoencoder.py:113-116).
@torch.inference_mode()
def vae_fixtures():
    import json
    sys.path.insert(0, "/root/reference/latent-diffusion")
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):          # the reference prints shapes while constructing
        from ldm.modules.diffusionmodules.model import Decoder
        cfgs = {
            # the reference's own first-stage config (latent-diffusion/train/configs/VAE_cifar.yaml): 3x16x16 latents -> 3x32x32
            "vae_decode_cifar": (dict(double_z=False, z_channels=3, resolution=32, in_channels=3, out_ch=3, ch=64, ch_mult=[1, 2],
                                      num_res_blocks=2, attn_resolutions=[], dropout=0.0), 3, (2, 3, 16, 16)),
            # attention inside an up level and three levels (exercises AttnBlock at 64 channels and two upsamplings)
            "vae_decode_attn": (dict(double_z=False, z_channels=4, resolution=32, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2],
                                     num_res_blocks=1, attn_resolutions=[16], dropout=0.0), 4, (1, 4, 8, 8)),
        }
        manifest = {}
        for name, (ddconfig, embed_dim, zshape) in cfgs.items():
            dec = Decoder(**ddconfig).eval()
            pq = torch.nn.Conv2d(embed_dim, ddconfig["z_channels"], 1).eval()
            from oracle.weights import synth_state_dict
            shapes = {"decoder." + k: tuple(v.shape) for k, v in dec.state_dict().items()}
            shapes.update({"post_quant_conv." + k: tuple(v.shape) for k, v in pq.state_dict().items()})
            sd = synth_state_dict(shapes, 31)
            dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
            pq.load_state_dict({k[len("post_quant_conv."):]: v for k, v in sd.items() if k.startswith("post_quant_conv.")}, strict=True)
            z = rnd(zshape, 300)
            y = dec(pq(z))
            manifest[name] = dict(ddconfig=ddconfig, embed_dim=embed_dim, shapes={k: list(s) for k, s in shapes.items()})
            save(name, z=z, y=y)
    with open(os.path.join(HERE, "manifest_vae.json"), "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    if "vae" not in sys.argv[1:]:
        main()
    vae_fixtures()
