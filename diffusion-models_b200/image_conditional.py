"""Image-conditional variant (reference: denoising_diffusion_image_conditional.py, `ic` below).

`Unet(*a, cond_channels=k)` widens the stem and concatenates `cond` on channels (ic:31-55) -- here the stem kernel
simply reads a second fp32 tensor, so no concatenated copy is made.  `ImageConditionalDenoisingDiffusion` threads
`cond` through the samplers (ic:78-223).

Deliberate differences from the reference, both documented in SURVEY.md section 0.6 / 7.2(8):
  * `sample()` under DDIM works (upstream passes `return_condition_image` into `sampling_timesteps`, ic:229, and runs
    zero U-Net steps);
  * the samplers take the condition as a keyword (`cond=`) in addition to drawing it from `condition_data_folder`.
"""
from __future__ import annotations

import random
from pathlib import Path

import torch

from .diffusion import DenoisingDiffusion, _KIND_DDIM, _KIND_DDPM
from .unet import Unet as _BaseUnet


class Unet(_BaseUnet):
    def __init__(self, *unet_args, cond_channels=0, **unet_kwargs):
        self.cond_channels = cond_channels
        super().__init__(*unet_args, cond_channels=cond_channels, **unet_kwargs)

    @torch.no_grad()
    def forward(self, x, time, *, cond=None, x_self_cond=None):
        """ic:51-55."""
        if cond is None and self.cond_channels:
            raise ValueError("cond is required: init_conv was built for channels + cond_channels inputs")
        b, _, h, w = x.shape
        eng = self.engine(b, h, w, device=x.device)
        self._stage_inputs(eng, x, time, x_self_cond, cond=cond)
        eng.run_time_path()
        eng.run_body()
        return eng.out.clone()


class ImageConditionalDenoisingDiffusion(DenoisingDiffusion):
    def __init__(self, *args, condition_data_folder=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.condition_data_folder = condition_data_folder

    def get_random_condition(self, batch, device):
        """ic:123-153 -- random images from `condition_data_folder`, resized/cropped to image_size, in [0,1]."""
        from PIL import Image
        from torchvision import transforms as T
        tf = T.Compose([T.Resize(self.image_size), T.CenterCrop(self.image_size), T.ToTensor()])
        paths = list(Path(self.condition_data_folder).glob("*.*"))
        picks = random.choices(paths, k=batch)
        return torch.stack([tf(Image.open(p).convert("RGB")) for p in picks], dim=0).to(device)

    def _prepare_cond(self, cond, batch):
        """Hook: LDM subclasses encode the condition image into latent space here (once, not per step)."""
        return cond

    @torch.no_grad()
    def p_sample_loop(self, shape, return_condition_image=False, return_all_timesteps=False, *, cond=None, noise=None,
                      step_noise=None, use_graph=True, trace=None):
        """ic:155-179."""
        raw = cond if cond is not None else self.get_random_condition(shape[0], self.device)
        times = list(reversed(range(self.num_timesteps)))
        ret = self._run_loop(_KIND_DDPM, tuple(shape), times, self._ddpm_coefs(times), x_T=noise, step_noise=step_noise,
                             return_all_timesteps=return_all_timesteps, use_graph=use_graph,
                             cond=self._prepare_cond(raw, shape[0]), trace=trace)
        return (raw, ret) if return_condition_image else ret

    @torch.no_grad()
    def ddim_sample(self, shape, sampling_timesteps=None, cond=None, return_all_timesteps=False, *, noise=None,
                    step_noise=None, use_graph=True, return_condition_image=False, trace=None):
        """ic:181-223 (same positional order: shape, sampling_timesteps, cond, return_all_timesteps)."""
        S = self.sampling_timesteps if sampling_timesteps is None else sampling_timesteps
        raw = cond if cond is not None else self.get_random_condition(shape[0], self.device)
        pairs = self._ddim_pairs(S)
        ret = self._run_loop(_KIND_DDIM, tuple(shape), [t for t, _ in pairs], self._ddim_coefs(pairs, self.ddim_sampling_eta),
                             x_T=noise, step_noise=step_noise, return_all_timesteps=return_all_timesteps,
                             use_graph=use_graph, cond=self._prepare_cond(raw, shape[0]), trace=trace)
        return (raw, ret) if return_condition_image else ret

    @torch.no_grad()
    def sample(self, batch_size=16, return_condition_image=False, return_all_timesteps=False, **kw):
        (h, w), channels = self.image_size, self.channels
        shape = (batch_size, channels, h, w)
        if self.is_ddim_sampling:
            return self.ddim_sample(shape, return_all_timesteps=return_all_timesteps,
                                    return_condition_image=return_condition_image, **kw)
        return self.p_sample_loop(shape, return_condition_image, return_all_timesteps=return_all_timesteps, **kw)

    @torch.no_grad()
    def interpolate(self, x1, x2, t=None, cond=None, lam=0.5, **kw):
        """ic:232-249 (same positional order: x1, x2, t, cond, lam)."""
        if cond is None:
            raise ValueError("interpolate() of an image-conditional model needs cond")
        return super().interpolate(x1, x2, t, lam, cond=self._prepare_cond(cond, x1.shape[0]), **kw)

    @torch.no_grad()
    def p_sample(self, x, t: int, cond=None, x_self_cond=None):
        """ic:114-121."""
        return super().p_sample(x, t, x_self_cond, cond=cond)

    @torch.no_grad()
    def model_predictions(self, x, t, cond=None, x_self_cond=None, clip_x_start=False, rederive_pred_noise=False):
        """ic:78-101 (cond is the third positional argument upstream)."""
        return super().model_predictions(x, t, x_self_cond, clip_x_start, rederive_pred_noise, cond=cond)

    @torch.no_grad()
    def p_mean_variance(self, x, t, cond=None, x_self_cond=None, clip_denoised=True):
        """ic:104-112."""
        return super().p_mean_variance(x, t, x_self_cond, clip_denoised, cond=cond)
