"""CPU oracle for the denoising hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain fp32 restatement (torch CPU tensor ops, no nn.Module
state) of the reference algorithm for the path BASELINE.json names: the
U-Net eps-prediction forward (`denoising_diffusion.py:349-390`) evaluated inside
the DDPM / DDIM sampling loops (`denoising_diffusion.py:647-708`).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import it -- as the checker or as the timed
CPU baseline, never as the product.  The product (`diffusion-models_b200/`)
never imports this package and has no CPU fallback.

Pinning: the reference ships no golden vectors, so the oracle is pinned
against outputs of the reference itself, imported unmodified from
/root/reference in the build container by `tests/golden/make_golden.py`; the
resulting fixtures are committed under `tests/golden/` and checked by
`tests/test_oracle_golden.py`.
"""
from .unet_ref import unet_forward, UnetConfig, infer_config  # noqa: F401
from .sampler_ref import (  # noqa: F401
    Schedule, make_schedule, ddim_time_pairs, ddim_sample, p_sample_loop,
    model_predictions, ddim_update, ddpm_update, q_sample, interpolate, ddpm_update_learned, p_sample_loop_learned,
    ddim_sample_guided,
)
from .vae_ref import vae_decode, decoder_forward  # noqa: F401
from .weights import synth_state_dict  # noqa: F401
