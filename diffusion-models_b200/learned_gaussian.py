"""`LearnedGaussianDiffusion` (reference `denoising_diffusion/learned_gaussian_diffusion.py:60-111`): ancestral sampling
with the network's second output half interpolating the log-variance between the posterior variance and beta_t
(Nichol & Dhariwal).  Sampling only; the hybrid eps + variational-bound loss (`p_losses`, lgd:113-150) is training and
out of scope.  The U-Net is `Unet(..., learned_variance=True)` (out_dim = 2 * channels); its forward runs on the same
engine, and the per-step update is `ddm_sampler_step_learned` (include/ddm_b200.h)."""
from __future__ import annotations

from typing import Sequence

import torch

from .diffusion import DenoisingDiffusion, _KIND_DDPM_LEARNED


class LearnedGaussianDiffusion(DenoisingDiffusion):
    def __init__(self, model, vb_loss_weight=0.001, *args, **kwargs):
        super().__init__(model, *args, **kwargs)
        assert model.out_dim == model.channels * 2, \
            "dimension out of unet must be twice the number of channels for learned variance - set learned_variance=True on the Unet"
        assert not model.self_condition, "not supported yet"        # lgd:71
        assert self.objective == "pred_noise", "learned-variance sampling is defined for pred_noise (lgd:95-104)"
        self.vb_loss_weight = vb_loss_weight

    def _learned_coefs(self, times: Sequence[int]) -> torch.Tensor:
        ra, rm1 = self.sqrt_recip_alphas_cumprod.cpu(), self.sqrt_recipm1_alphas_cumprod.cpu()
        c1, c2 = self.posterior_mean_coef1.cpu(), self.posterior_mean_coef2.cpu()
        min_log, max_log = self.posterior_log_variance_clipped.cpu(), torch.log(self.betas).cpu()     # lgd:95-96
        rows = torch.zeros((len(times), 8), dtype=torch.float32)
        for i, t in enumerate(times):
            rows[i, 0], rows[i, 1], rows[i, 2], rows[i, 3] = ra[t], rm1[t], c1[t], c2[t]
            rows[i, 4] = 1.0 if t > 0 else 0.0                                                      # dd:643
            rows[i, 5], rows[i, 6] = min_log[t], max_log[t]
        return rows

    @torch.no_grad()
    def p_sample_loop(self, shape, return_all_timesteps=False, *, noise=None, step_noise=None, use_graph=True, trace=None):
        """dd:647-664 with lgd:91-111 as p_mean_variance."""
        times = list(reversed(range(self.num_timesteps)))
        return self._run_loop(_KIND_DDPM_LEARNED, tuple(shape), times, self._learned_coefs(times), x_T=noise,
                              step_noise=step_noise, return_all_timesteps=return_all_timesteps, use_graph=use_graph, trace=trace)

    @torch.no_grad()
    def ddim_sample(self, *args, **kwargs):
        raise NotImplementedError("the reference's LearnedGaussianDiffusion.model_predictions (lgd:75-89) references undefined "
                                  "names and cannot run; only ancestral sampling is defined for this class")

    @torch.no_grad()
    def sample(self, batch_size=16, return_all_timesteps=False, **kw):
        (h, w), channels = self.image_size, self.channels
        return self.p_sample_loop((batch_size, channels, h, w), return_all_timesteps=return_all_timesteps, **kw)
