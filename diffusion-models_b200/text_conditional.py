"""Text-conditional variant (reference: denoising_diffusion_text_conditional.py, `tc` below).

Two mechanisms, chosen at construction like upstream (tc:86-125):
  * concat : text_proj(text) is fused with the time embedding through text_concat_proj (tc:146-152);
  * xattn  : three CrossAttention modules at the bottleneck width, each *replacing* x (tc:173-198).  Their
             to_k/to_v projections of the text are loop-invariant and are computed once per sampling call.

The upstream samplers take no text argument and read a pickle of CLIP embeddings from disk on every call
(tc:320-363).  Here they accept `text_emb=` as a keyword; without it they fall back to the same pickle lookup.
"""
from __future__ import annotations

import os
import pickle
import random
from pathlib import Path

import torch

from .diffusion import DenoisingDiffusion, _KIND_DDIM, _KIND_DDPM
from .unet import Unet as _BaseUnet


class Unet(_BaseUnet):
    def __init__(self, *, dim, init_dim=None, dim_mults=(1, 2, 4, 8), text_condition=True, text_emb_dim=512,
                 use_cross_attn=False, attn_dim_head=32, **base_kwargs):
        self.text_condition = text_condition
        self.use_cross_attn = use_cross_attn
        mode = None if not text_condition else ("xattn" if use_cross_attn else "concat")
        super().__init__(dim=dim, init_dim=init_dim, dim_mults=dim_mults, attn_dim_head=attn_dim_head, text_mode=mode,
                         text_emb_dim=text_emb_dim,
                         xattn_dim_head=attn_dim_head if isinstance(attn_dim_head, int) else attn_dim_head[-1], **base_kwargs)

    @torch.no_grad()
    def forward(self, x, time, text_emb=None, x_self_cond=None):
        """tc:131-214."""
        b, _, h, w = x.shape
        if text_emb is None or not self.text_condition:
            raise ValueError("text_emb is required (build the unconditional Unet for text-free sampling)")
        tokens = 0
        if self.use_cross_attn:
            tokens = 1 if text_emb.ndim == 2 else text_emb.shape[1]
        elif text_emb.dim() == 3 and text_emb.size(1) == 1:
            text_emb = text_emb.squeeze(1)                       # tc:147-148
        eng = self.engine(b, h, w, text_tokens=tokens, device=x.device)
        self._stage_inputs(eng, x, time, x_self_cond, text_emb=text_emb)
        if self.use_cross_attn:
            eng.run_text_path()
        eng.run_time_path()
        eng.run_body()
        return eng.out.clone()


class TextConditionalDenoisingDiffusion(DenoisingDiffusion):
    def __init__(self, *, model, embedding_file=None, **kwargs):
        super().__init__(model, **kwargs)
        if embedding_file is not None:
            assert os.path.exists(embedding_file), "Pre-computed caption embeddings file not found."
        self.embedding_file = Path(embedding_file) if embedding_file is not None else None

    def get_random_text_condition(self, batch, device):
        """tc:320-363 -- (embeddings [batch, dim], captions) drawn from the pickle of precomputed embeddings."""
        with open(self.embedding_file, "rb") as f:
            table = pickle.load(f)
        keys = random.choices(list(table.keys()), k=batch)
        embs, texts = [], []
        for key in keys:
            entry = table[key]
            i = random.randint(0, entry["embeddings"].shape[0] - 1)
            embs.append(torch.tensor(entry["embeddings"][i], dtype=torch.float))
            texts.append(entry["captions"][i])
        return torch.stack(embs, dim=0).to(device), texts

    def _text(self, batch, text_emb, save_path_for_text):
        if text_emb is not None:
            return text_emb
        text_emb, texts = self.get_random_text_condition(batch, self.device)
        if save_path_for_text is not None:                       # tc:376-380
            with open(save_path_for_text, "a" if os.path.exists(save_path_for_text) else "w") as f:
                for t in texts:
                    f.write(t + "\n")
        return text_emb

    @torch.no_grad()
    def p_sample_loop(self, shape, save_path_for_text=None, return_all_timesteps=False, *, text_emb=None, noise=None,
                      step_noise=None, use_graph=True, trace=None):
        """tc:366-393."""
        times = list(reversed(range(self.num_timesteps)))
        return self._run_loop(_KIND_DDPM, tuple(shape), times, self._ddpm_coefs(times), x_T=noise, step_noise=step_noise,
                              return_all_timesteps=return_all_timesteps, use_graph=use_graph,
                              text_emb=self._text(shape[0], text_emb, save_path_for_text), trace=trace)

    @torch.no_grad()
    def ddim_sample(self, shape, save_path_for_text=None, sampling_timesteps=None, return_all_timesteps=False, *,
                    text_emb=None, noise=None, step_noise=None, use_graph=True, trace=None):
        """tc:395-447."""
        S = self.sampling_timesteps if sampling_timesteps is None else sampling_timesteps
        pairs = self._ddim_pairs(S)
        return self._run_loop(_KIND_DDIM, tuple(shape), [t for t, _ in pairs], self._ddim_coefs(pairs, self.ddim_sampling_eta),
                              x_T=noise, step_noise=step_noise, return_all_timesteps=return_all_timesteps,
                              use_graph=use_graph, text_emb=self._text(shape[0], text_emb, save_path_for_text), trace=trace)

    @torch.no_grad()
    def sample(self, batch_size=16, save_path_for_text=None, return_all_timesteps=False, **kw):
        """tc:449-453."""
        (h, w), channels = self.image_size, self.channels
        fn = self.p_sample_loop if not self.is_ddim_sampling else self.ddim_sample
        return fn((batch_size, channels, h, w), save_path_for_text, return_all_timesteps=return_all_timesteps, **kw)

    @torch.no_grad()
    def interpolate(self, x1, x2, t=None, text_emb=None, lam=0.5, **kw):
        """tc:456-473 (same positional order: x1, x2, t, text_emb, lam)."""
        if text_emb is None:
            raise ValueError("interpolate() of a text-conditional model needs text_emb")
        return super().interpolate(x1, x2, t, lam, text_emb=text_emb, **kw)

    @torch.no_grad()
    def p_sample(self, x, t: int, text_emb=None, x_self_cond=None):
        """tc:309-316."""
        return super().p_sample(x, t, x_self_cond, text_emb=text_emb)

    @torch.no_grad()
    def model_predictions(self, x, t, text_emb=None, x_self_cond=None, clip_x_start=False, rederive_pred_noise=False):
        """tc:274-297."""
        return super().model_predictions(x, t, x_self_cond, clip_x_start, rederive_pred_noise, text_emb=text_emb)

    @torch.no_grad()
    def p_mean_variance(self, x, t, text_emb=None, x_self_cond=None, clip_denoised=True):
        """tc:299-307."""
        return super().p_mean_variance(x, t, x_self_cond, clip_denoised, text_emb=text_emb)
