"""Host-logic check on CPU: run the *real* UnetEngine plan (packed bf16 weights, ddm_conv_args structs, tap tables,
sub-pixel upsample phases, unshuffle view, scale/shift offsets, fused pre-norms) through tests/fake_lib.py and compare
the result with the fp32 oracle.  Differences are bf16 activation rounding only (rel-L2 of a few 1e-3)."""
import json
import os

import pytest
import torch

import diffusion_models_b200 as ddm
from diffusion_models_b200.engine import UnetEngine
from diffusion_models_b200.image_conditional import Unet as ImgUnet
from diffusion_models_b200.text_conditional import Unet as TextUnet
from oracle import unet_forward, infer_config, synth_state_dict
from fake_lib import FakeLib


def rel_l2(a, b):
    return ((a - b).norm() / b.norm()).item()


def run_case(model, seed, x, t, infer_kw=None, time_rows=None, **extra):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = synth_state_dict(shapes, seed)
    model.load_state_dict(sd)
    lib = FakeLib()
    B, _, H, W = x.shape
    text = extra.get("text_emb")
    tokens = 0
    if text is not None and model.spec.text_mode == "xattn":
        tokens = 1 if text.ndim == 2 else text.shape[1]
    eng = UnetEngine(model.spec, dict(model.named_parameters()), B, H, W, torch.device("cpu"),
                     time_rows=time_rows or B, text_tokens=tokens, lib=lib)
    lib.attach(eng)
    model._stage_inputs(eng, x, t, extra.get("x_self_cond"), cond=extra.get("cond"), text_emb=text)
    eng.run_text_path()
    eng.run_time_path()
    eng.run_body()
    cfg = infer_config(sd, **(infer_kw or {}))
    taps = {}
    with torch.inference_mode():
        ref = unet_forward(sd, x, t, cfg, taps=taps, **extra)
    return eng, ref, taps


def g(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def test_base_unet_32():
    eng, ref, taps = run_case(ddm.Unet(dim=64, dim_mults=(1, 2, 4, 8)), 0, g((2, 3, 32, 32), 1), torch.tensor([999, 17]))
    # per-layer: every named activation of the plan against the oracle's activation of the same name
    worst = 0.0
    for name, act in eng.taps.items():
        if name in taps:
            e = rel_l2(act.float().permute(0, 3, 1, 2), taps[name])
            worst = max(worst, e)
            assert e < 3e-2, (name, e)
    assert rel_l2(eng.out, ref) < 2e-2, (rel_l2(eng.out, ref), worst)


def test_small_unet_dim32_heads2():
    eng, ref, _ = run_case(ddm.Unet(dim=32, dim_mults=(1, 2), attn_heads=2, attn_dim_head=16), 5, g((2, 3, 16, 16), 3),
                           torch.tensor([3, 640]), infer_kw=dict(heads=2, dim_head=16))
    assert rel_l2(eng.out, ref) < 2e-2


def test_uniform_time_row():
    eng, ref, _ = run_case(ddm.Unet(dim=32, dim_mults=(1, 2)), 11, g((3, 3, 16, 16), 4), torch.tensor([321, 321, 321]), time_rows=1)
    assert rel_l2(eng.out, ref) < 2e-2


def test_self_condition():
    eng, ref, _ = run_case(ddm.Unet(dim=32, dim_mults=(1, 2), self_condition=True), 6, g((1, 3, 16, 16), 5), torch.tensor([250]),
                           infer_kw=dict(self_condition=True), x_self_cond=g((1, 3, 16, 16), 40))
    assert rel_l2(eng.out, ref) < 2e-2


def test_image_conditional():
    eng, ref, _ = run_case(ImgUnet(dim=32, dim_mults=(1, 2), channels=4, cond_channels=4), 7, g((2, 4, 16, 16), 6),
                           torch.tensor([777, 1]), cond=g((2, 4, 16, 16), 50))
    assert rel_l2(eng.out, ref) < 2e-2


def test_text_cross_attention():
    eng, ref, _ = run_case(TextUnet(dim=32, dim_mults=(1, 2), channels=4, text_condition=True, use_cross_attn=True), 8,
                           g((2, 4, 16, 16), 7), torch.tensor([123, 500]), text_emb=g((2, 5, 512), 60))
    assert rel_l2(eng.out, ref) < 2e-2


def test_text_concat():
    eng, ref, _ = run_case(TextUnet(dim=32, dim_mults=(1, 2), channels=4, text_condition=True, use_cross_attn=False), 9,
                           g((2, 4, 16, 16), 8), torch.tensor([5, 900]), text_emb=g((2, 512), 70))
    assert rel_l2(eng.out, ref) < 2e-2


def test_full_attention_everywhere_and_odd_sizes():
    eng, ref, _ = run_case(ddm.Unet(dim=32, dim_mults=(1, 2), full_attn=(True, True)), 10, g((1, 3, 8, 24), 9), torch.tensor([42]))
    assert rel_l2(eng.out, ref) < 2e-2


@pytest.mark.parametrize("name", ["vae_decode_cifar", "vae_decode_attn"])
def test_vae_decode_plan(name):
    """The real VaeDecodeEngine plan (post_quant as a 1x1 stem, padded conv_in, GroupNorm pre-passes, fused residuals, stacked
    q|k|v GEMM + single-head attention, sub-pixel upsampling) through the C-ABI stand-in, against the fp32 oracle."""
    import numpy as np
    from conftest import GOLDEN
    from oracle import vae_decode
    with open(os.path.join(GOLDEN, "manifest_vae.json")) as f:
        m = json.load(f)[name]
    vae = ddm.VQDecoder(ddconfig=m["ddconfig"], embed_dim=m["embed_dim"])
    assert {k: list(v.shape) for k, v in vae.state_dict().items()} == m["shapes"]
    sd = synth_state_dict({k: tuple(v) for k, v in m["shapes"].items()}, 31)
    vae.load_state_dict(sd)
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        zin, want = torch.from_numpy(z["z"]), torch.from_numpy(z["y"])
    lib = FakeLib()
    eng = vae.engine(zin.shape[0], zin.shape[2], zin.shape[3], torch.device("cpu"), lib=lib)
    lib.attach(eng)
    eng.z.copy_(zin)
    eng.run()
    with torch.inference_mode():
        ref = vae_decode(sd, zin)
    assert rel_l2(eng.out, ref) < 2e-2 and rel_l2(eng.out, want) < 2e-2, (rel_l2(eng.out, ref), rel_l2(eng.out, want))
