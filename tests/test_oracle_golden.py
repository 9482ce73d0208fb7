"""Pin the oracle (oracle/) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only.  fp32 on both sides: the tolerance only
absorbs reassociation inside ATen kernels (conv algorithm choice, einsum order)."""
import json
import os

import pytest
import torch

from oracle import (unet_forward, infer_config, synth_state_dict, make_schedule, ddim_time_pairs,
                    ddim_sample, p_sample_loop, interpolate, p_sample_loop_learned, ddim_sample_guided)
from conftest import GOLDEN

with open(os.path.join(GOLDEN, "manifest.json")) as f:
    MANIFEST = json.load(f)
with open(os.path.join(GOLDEN, "manifest_extra.json")) as f:          # tests/golden/make_golden_extra.py
    MANIFEST.update(json.load(f))

UNET_CASES = {
    # name: (seed, infer_config kwargs)
    "unet_base_32": (0, {}),
    "unet_base_64": (0, {}),
    "unet_small_16": (5, dict(heads=2, dim_head=16)),
    "unet_selfcond_32": (6, dict(self_condition=True)),
    "unet_imgcond_32": (7, {}),
    "unet_text_xattn_32": (8, {}),
    "unet_text_concat_32": (9, {}),
    "unet_full_attn_all_16": (10, {}),
    # BASELINE shapes (tests/golden/make_golden_r2.py): C3 image-conditional latents, C4 text cross-attention, 128 px
    "unet_imgcond_64": (7, {}),
    "unet_text_xattn_64": (8, {}),
    "unet_base_128": (0, {}),
}
MANIFEST_OF = {"unet_imgcond_64": "unet_imgcond_32", "unet_text_xattn_64": "unet_text_xattn_32", "unet_base_128": "unet_base_32"}


def _close(a, b, rtol=2e-4):
    scale = b.abs().max().item()
    err = (a - b).abs().max().item()
    assert err <= rtol * max(scale, 1.0), f"max-abs err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("name", sorted(UNET_CASES))
def test_unet_forward_matches_reference(name, golden):
    seed, kw = UNET_CASES[name]
    g = golden(name)
    sd = synth_state_dict(MANIFEST[MANIFEST_OF.get(name, name)], seed)
    cfg = infer_config(sd, **kw)
    extra = {k: g[k] for k in ("x_self_cond", "cond", "text_emb") if k in g}
    with torch.inference_mode():
        y = unet_forward(sd, g["x"], g["t"], cfg, **extra)
    _close(y, g["y"])


def test_schedules_bit_exact(golden):
    g = golden("schedules")
    for kind, T in (("linear", 1000), ("cosine", 1000), ("sigmoid", 1000), ("cosine", 6)):
        s = make_schedule(T, kind)
        for field in s.__dataclass_fields__:
            assert torch.equal(getattr(s, field), g[f"{kind}_{T}_{field}"]), (kind, T, field)


def test_ddim_time_pairs(golden):
    g = golden("ddim_pairs")
    for key in g:
        T, S = (int(v[1:]) for v in key.split("_"))
        assert ddim_time_pairs(T, S) == [tuple(r) for r in g[key].tolist()]


def _model(name="unet_base_32", seed=0, **fw):
    sd = synth_state_dict(MANIFEST[name], seed)
    cfg = infer_config(sd)
    return lambda x, t, sc: unet_forward(sd, x, t, cfg, x_self_cond=sc, **fw)


@torch.inference_mode()
def test_ddim_eta0(golden):
    g = golden("ddim_eta0_S5")
    y = ddim_sample(_model(), make_schedule(1000), g["x_T"], 5, eta=0.0, noises=list(g["noises"]))
    _close(y, g["y"], 1e-3)


@torch.inference_mode()
def test_ddim_eta1_all_timesteps(golden):
    g = golden("ddim_eta1_S4")
    y = ddim_sample(_model(), make_schedule(1000), g["x_T"], 4, eta=1.0, noises=list(g["noises"]),
                    return_all_timesteps=True)
    assert y.shape == g["y"].shape
    _close(y, g["y"], 1e-3)


@torch.inference_mode()
def test_ddpm_loop(golden):
    g = golden("ddpm_T6")
    y = p_sample_loop(_model(), make_schedule(6, "cosine"), g["x_T"], noises=list(g["noises"]))
    _close(y, g["y"], 1e-3)


@torch.inference_mode()
def test_learned_variance(golden):
    # Unet(learned_variance=True): 2C output channels; LearnedGaussianDiffusion ancestral loop (lgd:91-111)
    sd = synth_state_dict(MANIFEST["unet_learned_var_32"], 12)
    cfg = infer_config(sd)
    g = golden("unet_learned_var_32")
    y = unet_forward(sd, g["x"], g["t"], cfg)
    assert y.shape == (2, 6, 32, 32)
    _close(y, g["y"])
    g = golden("learned_var_T6")
    model = lambda x, t, sc: unet_forward(sd, x, t, cfg)
    y = p_sample_loop_learned(model, make_schedule(6, "cosine"), g["x_T"], noises=list(g["noises"]))
    _close(y, g["y"], 1e-3)


@torch.inference_mode()
def test_interpolate(golden):
    g = golden("interpolate_T6")          # dd:785-803: q_sample both, blend, ancestral steps t-1 .. 0, raw output
    y = interpolate(_model(), make_schedule(6, "cosine"), g["x1"], g["x2"], t=4, lam=0.3, q_noise=list(g["q_noise"]),
                    noises=list(g["noises"]))
    _close(y, g["y"], 1e-3)


@torch.inference_mode()
def test_ddim_pred_v_cosine(golden):
    g = golden("ddim_predv_S3")
    y = ddim_sample(_model(), make_schedule(1000, "cosine"), g["x_T"], 3, objective="pred_v")
    _close(y, g["y"], 1e-3)


@torch.inference_mode()
def test_sample_dispatch(golden):
    g = golden("sample_dispatch_S3")          # sample() -> ddim_sample when S < T (dd:779-783)
    y = ddim_sample(_model(), make_schedule(1000), g["x_T"], 3)
    _close(y, g["y"], 1e-3)


@torch.inference_mode()
def test_ddim_image_conditional(golden):
    g = golden("ddim_imgcond_S3")
    y = ddim_sample(_model("unet_imgcond_32", 7, cond=g["cond"]), make_schedule(1000), g["x_T"], 3, unnormalize=False)
    _close(y, g["y"], 1e-3)


@torch.inference_mode()
def test_ddim_text_cross_attention(golden):
    g = golden("ddim_text_xattn_S3")
    y = ddim_sample(_model("unet_text_xattn_32", 8, text_emb=g["text_emb"]), make_schedule(1000), g["x_T"], 3,
                    unnormalize=False)
    _close(y, g["y"], 1e-3)


@torch.inference_mode()
def test_ddim_guided(golden):
    g = golden("ddim_guided_S4")          # dd:710-777 with guide / mask, eta = 0.5
    y = ddim_sample_guided(_model(), make_schedule(1000), g["x_T"], 4, eta=0.5, guide=g["guide"], mask=g["mask"],
                           noises=list(g["noises"]), guide_noises=list(g["guide_noises"]))
    _close(y, g["y"], 1e-3)
    g = golden("ddim_guided_noguide_S4")  # no guide, clip_denoised=False: raw eps, unclamped x0
    y = ddim_sample_guided(_model(), make_schedule(1000), g["x_T"], 4, eta=0.5, clip_denoised=False, noises=list(g["noises"]))
    _close(y, g["y"], 1e-3)


@pytest.mark.parametrize("name", ["vae_decode_cifar", "vae_decode_attn"])
@torch.inference_mode()
def test_vae_decode_matches_reference(name, golden):
    """oracle/vae_ref.py against the reference's Decoder + post_quant_conv (tests/golden/make_golden_r2.py vae_fixtures)."""
    from oracle import vae_decode
    with open(os.path.join(GOLDEN, "manifest_vae.json")) as f:
        m = json.load(f)[name]
    sd = synth_state_dict({k: tuple(v) for k, v in m["shapes"].items()}, 31)
    g = golden(name)
    _close(vae_decode(sd, g["z"]), g["y"])
