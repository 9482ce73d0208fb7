"""Latent-diffusion wrappers (reference: latent-diffusion/ldm/models/latent_diffusion*.py).

The LDM classes are the same samplers run on VAE latents: `normalize`/`unnormalize` are identity (x0 is still clamped
to [-1, 1] in latent space), and `sample()` decodes the final latents through `vae.decode`
(latent_diffusion.py:25-26,60-67).  The step loop is the B200 hot path; the VAE is any object with `.decode()` /
`.encode()` (the reference's VQModel) and runs once per call, after the loop -- SURVEY.md section 8(f) "next" #1.
"""
from __future__ import annotations

import torch

from .diffusion import DenoisingDiffusion, identity
from .image_conditional import ImageConditionalDenoisingDiffusion
from .text_conditional import TextConditionalDenoisingDiffusion


class _LatentMixin:
    def _init_latent(self, vae, latent_shape, cond_vae=None):
        self.vae = vae
        self.cond_vae = cond_vae if cond_vae is not None else vae
        self.latent_channels = latent_shape[0]
        self.model.channels = self.latent_channels
        self.normalize = identity
        self.unnormalize = identity
        self._auto_normalize = False
        for v in {id(self.vae): self.vae, id(self.cond_vae): self.cond_vae}.values():
            if hasattr(v, "eval"):
                v.eval()
            if hasattr(v, "parameters"):
                for p in v.parameters():
                    p.requires_grad = False

    def encode(self, images, cond=False):
        """latent_diffusion.py:35-41 -- VQModel.encode may return (latents, ...)."""
        with torch.no_grad():
            z = (self.cond_vae if cond else self.vae).encode(images)
        return z[0] if isinstance(z, tuple) else z

    def decode(self, latents):
        """latent_diffusion.py:44-48."""
        with torch.no_grad():
            return self.vae.decode(latents)


class LatentDiffusion(_LatentMixin, DenoisingDiffusion):
    def __init__(self, model, vae, latent_shape, **kwargs):
        kwargs.setdefault("auto_normalize", False)
        super().__init__(model, image_size=latent_shape[1], **kwargs)
        self._init_latent(vae, latent_shape)

    @torch.no_grad()
    def sample(self, batch_size=16, return_all_timesteps=False, **kw):
        """latent_diffusion.py:60-67."""
        return self.decode(super().sample(batch_size, return_all_timesteps, **kw))


class ImageConditionalLatentDiffusion(_LatentMixin, ImageConditionalDenoisingDiffusion):
    def __init__(self, model, vae, latent_shape, init_image_size=None, cond_vae=None, **kwargs):
        kwargs.setdefault("auto_normalize", False)
        super().__init__(model, image_size=latent_shape[1], **kwargs)
        self._init_latent(vae, latent_shape, cond_vae)
        self.init_image_size = init_image_size

    def _prepare_cond(self, cond, batch):
        """The reference re-encodes the condition image on every one of the 1000 steps
        (latent_diffusion_image_conditional.py:128); it is loop-invariant, so encode once.  A tensor that already has
        the latent shape is passed through (upstream's inherited ddim_sample expects a latent)."""
        if cond.shape[-2:] == tuple(self.image_size) and cond.shape[1] == self.model.spec.cond_channels:
            return cond
        return self.encode(cond, cond=True)

    @torch.no_grad()
    def sample(self, batch_size=16, return_condition_image=False, return_all_timesteps=False, **kw):
        """latent_diffusion_image_conditional.py:143-168."""
        res = super().sample(batch_size, return_condition_image, return_all_timesteps, **kw)
        if return_condition_image:
            return res[0], self.decode(res[1])
        return self.decode(res)


class TextConditionalLatentDiffusion(_LatentMixin, TextConditionalDenoisingDiffusion):
    """Upstream's constructor always raises (it calls a keyword-only ctor positionally); this one
    keeps the documented signature and works."""

    def __init__(self, model, vae, latent_shape, embedding_file=None, **kwargs):
        kwargs.setdefault("auto_normalize", False)
        super().__init__(model=model, embedding_file=embedding_file, image_size=latent_shape[1], **kwargs)
        self._init_latent(vae, latent_shape)

    @torch.no_grad()
    def sample(self, batch_size=16, save_path_for_text=None, return_all_timesteps=False, **kw):
        """latent_diffusion_text_conditional.py:80-100."""
        return self.decode(super().sample(batch_size, save_path_for_text, return_all_timesteps, **kw))
