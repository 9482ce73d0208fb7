"""`DDIMSampler` -- the CompVis-style call shape BASELINE.json's north star names.  The reference has no such class
(its LDM wrappers inherit `ddim_sample`, latent_diffusion.py:60-67); this shim maps `sample(S, batch_size, shape,
eta=..., x_T=...)` onto `DenoisingDiffusion.ddim_sample` so code written against either surface runs on the B200 path.
"""
from __future__ import annotations

import torch


class DDIMSampler:
    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model               # a DenoisingDiffusion (or subclass) from this package

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, eta=0., x_T=None, verbose=False, **kwargs):
        """shape = (C, H, W).  Returns (samples, intermediates) like CompVis' sampler; intermediates is an empty dict."""
        old_eta = self.model.ddim_sampling_eta
        self.model.ddim_sampling_eta = eta
        try:
            full = (batch_size,) + tuple(shape)
            kw = dict(noise=x_T)
            if conditioning is not None:
                name = "cond" if hasattr(self.model, "condition_data_folder") else "text_emb"
                kw[name] = conditioning
            out = self.model.ddim_sample(full, sampling_timesteps=S, **kw)
        finally:
            self.model.ddim_sampling_eta = old_eta
        return out, {}
