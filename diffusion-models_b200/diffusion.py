"""`DenoisingDiffusion` -- drop-in for the reference sampler class (denoising_diffusion.py:435-803; exported as
`GaussianDiffusion` too) running the step loop on the B200 kernels.

Per step the reference launches ~3100 ATen ops; here a step is ~110 launches of our own kernels (U-Net plan +
one fused posterior-update kernel), captured once in a CUDA graph and replayed for every timestep with a
device-side step counter selecting the per-step coefficients and the per-step scale/shift row.
"""
from __future__ import annotations

import math
from collections import namedtuple
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .engine import UnetEngine

ModelPrediction = namedtuple("ModelPrediction", ["pred_noise", "pred_x_start"])

_OBJECTIVES = {"pred_noise": 0, "pred_x0": 1, "pred_v": 2}
MAX_GRAPH_STEPS = 250       # steps captured into one CUDA graph (x ~105 kernel nodes each)


def _p(t):
    return None if t is None else t.data_ptr()


def identity(t, *args, **kwargs):
    return t


def normalize_to_neg_one_to_one(img):        # utils.py:45-46
    return img * 2 - 1


def unnormalize_to_zero_to_one(t):           # utils.py:48-49
    return (t + 1) * 0.5


def extract(a, t, x_shape):                  # dd:394-397
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def linear_beta_schedule(timesteps):         # dd:399-406
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def cosine_beta_schedule(timesteps, s=0.008):   # dd:408-418
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps
    ac = torch.cos((x + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def sigmoid_beta_schedule(timesteps, start=-3, end=3, tau=1, clamp_min=1e-5):   # dd:420-433
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps
    v_start = torch.tensor(start / tau).sigmoid()
    v_end = torch.tensor(end / tau).sigmoid()
    ac = (-((x * (end - start) + start) / tau).sigmoid() + v_end) / (v_end - v_start)
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


class DenoisingDiffusion(nn.Module):
    def __init__(self, model, *, image_size, timesteps=1000, sampling_timesteps=None, objective="pred_noise",
                 beta_schedule="linear", schedule_fn_kwargs=dict(), ddim_sampling_eta=0., auto_normalize=True,
                 offset_noise_strength=0., min_snr_loss_weight=False, min_snr_gamma=5, immiscible=False, ddpm=True,
                 hybrid_loss=False):
        super().__init__()
        assert not (type(self) == DenoisingDiffusion and model.channels != model.out_dim)
        assert not getattr(model, "random_or_learned_sinusoidal_cond", False)
        self.model = model
        self.channels = model.channels
        self.self_condition = model.self_condition
        if isinstance(image_size, int):
            image_size = (image_size, image_size)
        assert isinstance(image_size, (tuple, list)) and len(image_size) == 2, \
            "image size must be a integer or a tuple/list of two integers"
        self.image_size = tuple(image_size)
        assert objective in _OBJECTIVES, "objective must be either pred_noise, pred_x0 or pred_v"
        self.objective = objective

        fns = {"linear": linear_beta_schedule, "cosine": cosine_beta_schedule, "sigmoid": sigmoid_beta_schedule}
        if beta_schedule not in fns:
            raise ValueError(f"unknown beta schedule {beta_schedule}")
        betas = fns[beta_schedule](timesteps, **schedule_fn_kwargs)
        alphas = 1. - betas
        alphas_cumprod = torch.cumprod(alphas, dim=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.)
        self.num_timesteps = int(betas.shape[0])
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else self.num_timesteps
        assert self.sampling_timesteps <= self.num_timesteps
        self.is_ddim_sampling = self.sampling_timesteps < self.num_timesteps
        self.ddim_sampling_eta = ddim_sampling_eta

        reg = lambda name, val: self.register_buffer(name, val.to(torch.float32))   # fp64 -> fp32 like dd:498-500
        reg("betas", betas)
        reg("alphas_cumprod", alphas_cumprod)
        reg("alphas_cumprod_prev", alphas_cumprod_prev)
        reg("sqrt_alphas_cumprod", torch.sqrt(alphas_cumprod))
        reg("sqrt_one_minus_alphas_cumprod", torch.sqrt(1. - alphas_cumprod))
        reg("log_one_minus_alphas_cumprod", torch.log(1. - alphas_cumprod))
        reg("sqrt_recip_alphas_cumprod", torch.sqrt(1. / alphas_cumprod))
        reg("sqrt_recipm1_alphas_cumprod", torch.sqrt(1. / alphas_cumprod - 1))
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        reg("posterior_variance", posterior_variance)
        reg("posterior_log_variance_clipped", torch.log(posterior_variance.clamp(min=1e-20)))
        reg("posterior_mean_coef1", betas * torch.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod))
        reg("posterior_mean_coef2", (1. - alphas_cumprod_prev) * torch.sqrt(alphas) / (1. - alphas_cumprod))
        # training-only knobs are accepted for signature parity; loss_weight kept so reference state_dicts load strictly
        self.immiscible, self.offset_noise_strength, self.hybrid_loss = immiscible, offset_noise_strength, hybrid_loss
        if ddpm:
            reg("loss_weight", torch.ones(timesteps, dtype=torch.float32))
        else:
            snr = alphas_cumprod / (1 - alphas_cumprod)
            clipped = snr.clone()
            if min_snr_loss_weight:
                clipped.clamp_(max=min_snr_gamma)
            reg("loss_weight", {"pred_noise": clipped / snr, "pred_x0": clipped, "pred_v": clipped / (snr + 1)}[objective])
        self.normalize = normalize_to_neg_one_to_one if auto_normalize else identity
        self.unnormalize = unnormalize_to_zero_to_one if auto_normalize else identity
        self._auto_normalize = auto_normalize
        self._seed_calls = 0

    @property
    def device(self):
        return self.betas.device

    # ------------------------------------------------------------------ small reference-API helpers (dd:570-601)
    def predict_start_from_noise(self, x_t, t, noise):
        return extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - \
            extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise

    def predict_noise_from_start(self, x_t, t, x0):
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - x0) / \
            extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape)

    def predict_v(self, x_start, t, noise):
        return extract(self.sqrt_alphas_cumprod, t, x_start.shape) * noise - \
            extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * x_start

    def predict_start_from_v(self, x_t, t, v):
        return extract(self.sqrt_alphas_cumprod, t, x_t.shape) * x_t - \
            extract(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * v

    def q_posterior(self, x_start, x_t, t):
        mean = extract(self.posterior_mean_coef1, t, x_t.shape) * x_start + extract(self.posterior_mean_coef2, t, x_t.shape) * x_t
        return mean, extract(self.posterior_variance, t, x_t.shape), extract(self.posterior_log_variance_clipped, t, x_t.shape)

    def _model_kwargs(self, **cond):
        return {k: v for k, v in cond.items() if v is not None}

    @torch.no_grad()
    def model_predictions(self, x, t, x_self_cond=None, clip_x_start=False, rederive_pred_noise=False, **cond):
        """dd:603-626.  Network on our kernels; the few-byte elementwise tail of this *compatibility* method uses
        tensor ops (the sampling loops below use the fused K9 kernel instead)."""
        kw = self._model_kwargs(**cond)
        out = self.model(x, t, x_self_cond=x_self_cond, **kw) if kw else self.model(x, t, x_self_cond)
        clip = (lambda z: z.clamp(-1., 1.)) if clip_x_start else identity
        if self.objective == "pred_noise":
            pred_noise = out
            x_start = clip(self.predict_start_from_noise(x, t, pred_noise))
            if clip_x_start and rederive_pred_noise:
                pred_noise = self.predict_noise_from_start(x, t, x_start)
        elif self.objective == "pred_x0":
            x_start = clip(out)
            pred_noise = self.predict_noise_from_start(x, t, x_start)
        else:
            x_start = clip(self.predict_start_from_v(x, t, out))
            pred_noise = self.predict_noise_from_start(x, t, x_start)
        return ModelPrediction(pred_noise, x_start)

    @torch.no_grad()
    def p_mean_variance(self, x, t, x_self_cond=None, clip_denoised=True, **cond):
        preds = self.model_predictions(x, t, x_self_cond=x_self_cond, **cond)
        x_start = preds.pred_x_start
        if clip_denoised:
            x_start.clamp_(-1., 1.)
        mean, var, logvar = self.q_posterior(x_start=x_start, x_t=x, t=t)
        return mean, var, logvar, x_start

    # ------------------------------------------------------------------ step tables (host, fp32 like the reference)
    def _ddim_pairs(self, S: int) -> List[Tuple[int, int]]:
        times = torch.linspace(-1, self.num_timesteps - 1, steps=S + 1)           # dd:672
        times = list(reversed(times.int().tolist()))
        return list(zip(times[:-1], times[1:]))

    def _ddim_coefs(self, pairs: Sequence[Tuple[int, int]], eta: float) -> torch.Tensor:
        ac = self.alphas_cumprod.detach().cpu()
        ra, rm1 = self.sqrt_recip_alphas_cumprod.cpu(), self.sqrt_recipm1_alphas_cumprod.cpu()
        sac, s1m = self.sqrt_alphas_cumprod.cpu(), self.sqrt_one_minus_alphas_cumprod.cpu()
        rows = torch.zeros((len(pairs), 8), dtype=torch.float32)
        for i, (t, tn) in enumerate(pairs):
            rows[i, 0], rows[i, 1], rows[i, 6], rows[i, 7] = ra[t], rm1[t], sac[t], s1m[t]
            if tn < 0:
                rows[i, 5] = 1.0                                                 # dd:686-689 img = x_start
                continue
            alpha, alpha_next = ac[t], ac[tn]                                    # dd:691-695, 0-d fp32 tensor math
            sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            c = (1 - alpha_next - sigma ** 2).sqrt()
            rows[i, 2], rows[i, 3], rows[i, 4] = alpha_next.sqrt(), c, sigma
        return rows

    def _ddpm_coefs(self, times: Sequence[int]) -> torch.Tensor:
        ra, rm1 = self.sqrt_recip_alphas_cumprod.cpu(), self.sqrt_recipm1_alphas_cumprod.cpu()
        sac, s1m = self.sqrt_alphas_cumprod.cpu(), self.sqrt_one_minus_alphas_cumprod.cpu()
        c1, c2 = self.posterior_mean_coef1.cpu(), self.posterior_mean_coef2.cpu()
        lv = self.posterior_log_variance_clipped.cpu()
        rows = torch.zeros((len(times), 8), dtype=torch.float32)
        for i, t in enumerate(times):
            rows[i, 0], rows[i, 1], rows[i, 2], rows[i, 3] = ra[t], rm1[t], c1[t], c2[t]
            rows[i, 4] = (0.5 * lv[t]).exp() if t > 0 else 0.0                  # dd:643-644
            rows[i, 6], rows[i, 7] = sac[t], s1m[t]
        return rows

    @staticmethod
    def _rank() -> int:
        return torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0

    def _next_seed(self) -> int:
        """Key of a cached loop's in-kernel Philox stream (a launch argument baked into the captured graph)."""
        self._seed_calls += 1
        return (torch.initial_seed() * 1000003 + self._seed_calls * 7919 + 17 * self._rank()) % (2 ** 63)

    @classmethod
    def _call_salt(cls) -> int:
        """One draw per sampling call from torch's (CPU) default generator: like the reference, a call is reproducible after
        `torch.manual_seed(s)` and differs from the previous call otherwise.  It seeds the call's x_T and, through the device
        counter's second slot, salts the step noise of a replayed graph; the rank is mixed in so that the ranks of a sharded
        run draw different samples under the same `manual_seed` (the usual torchrun pattern)."""
        base = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        return (base * 2654435761 + cls._rank() * 40503 + 1) % (2 ** 62)

    # ------------------------------------------------------------------ the fused loop
    @torch.no_grad()
    def _run_loop(self, kind: int, shape, times: Sequence[int], coefs: torch.Tensor, *, x_T=None, step_noise=None,
                  return_all_timesteps=False, use_graph=True, cond=None, text_emb=None, trace=None, raw=False, guided=None):
        """Shared DDIM / DDPM driver: [select scale-shift row -> U-Net plan -> fused update] per step."""
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("sampling runs only on a CUDA (B200) device; there is no CPU fallback")
        B, C, H, W = shape
        S = len(times)
        model = self.model
        per_sample_time = getattr(model.spec, "text_mode", None) == "concat" and text_emb is not None
        text_tokens = 0
        if text_emb is not None and model.spec.text_mode == "xattn":
            text_tokens = 1 if text_emb.ndim == 2 else text_emb.shape[1]
        eng: UnetEngine = model.engine(B, H, W, time_rows=B if per_sample_time else 1, text_tokens=text_tokens, device=dev)
        lib = eng.lib
        stream = torch.cuda.current_stream(dev)

        salt = self._call_salt()
        if x_T is None:                                                           # dd:651,676
            gen = torch.Generator(device=dev)
            gen.manual_seed(salt)
            x_T = torch.randn(shape, device=dev, generator=gen)
        # a conditional network must be given its condition: the engine buffers would otherwise still hold the previous
        # call's (or zeros), silently
        if eng.cond is not None and cond is None:
            raise ValueError("this model is image-conditional: pass cond=")
        if eng.text is not None and text_emb is None:
            raise ValueError("this model is text-conditional: pass text_emb=")
        if cond is not None:
            eng.cond.copy_(cond)
        if text_emb is not None and eng.text is not None:
            eng.text.copy_(text_emb.reshape(eng.text.shape))
            eng.run_text_path()
        if eng.x_self_cond is not None:
            eng.x_self_cond.zero_()
        obj = _OBJECTIVES[self.objective]
        numel = B * C * H * W
        eager = (not use_graph) or return_all_timesteps or per_sample_time or trace is not None or guided is not None

        # A loop = device tables (per-step coefficients, per-step scale/shift rows), a {step, epoch} counter and the
        # captured CUDA graph of one step.  It only depends on (engine, sampler kind, objective, timestep list,
        # coefficients), so it is built once and replayed by every later call: a sampling call then costs one copy of
        # x_T, S graph launches and the final unnormalise -- no per-call capture, table build or synchronisation.
        cacheable = not eager and step_noise is None
        key = (id(self), kind, obj, tuple(times))
        loop = eng.loops.get(key) if cacheable else None
        if loop is not None and not torch.equal(loop["coefs_host"], coefs):
            loop = None
        if loop is None:
            loop = dict(coefs_host=coefs.clone(), coef_dev=coefs.to(dev).contiguous(),
                        counter=torch.zeros((2,), dtype=torch.int32, device=dev), graph=None, per_step=0,
                        seed=self._next_seed(), epoch=0, ss_table=None)
            if not per_sample_time:      # per-step conditioning rows (time is batch-invariant, dd:641,682)
                tvals = torch.tensor([float(t) for t in times], dtype=torch.float32, device=dev)
                loop["ss_table"] = eng.build_step_table(tvals)
            if cacheable:
                eng.loops[key] = loop
        coef_dev, counter, ss_table, seed = loop["coef_dev"], loop["counter"], loop["ss_table"], loop["seed"]
        loop["epoch"] = salt & 0x7FFFFFFF
        counter.copy_(torch.tensor([0, loop["epoch"]], dtype=torch.int32))

        noise_ptr, noise_stride = None, 0
        if step_noise is not None:
            step_noise = step_noise.to(dev, torch.float32).contiguous()
            assert step_noise.shape[1:] == tuple(shape)
            pad = S - step_noise.shape[0]
            if pad > 0:                       # the reference draws no noise on the last step (dd:643, dd:686-689)
                step_noise = torch.cat([step_noise, torch.zeros((pad,) + tuple(shape), device=dev)], dim=0)
            noise_ptr, noise_stride = step_noise.data_ptr(), numel
        x0_ptr = eng.x_self_cond.data_ptr() if eng.x_self_cond is not None else None
        x0_buf = None
        if trace is not None and x0_ptr is None:
            x0_buf = torch.zeros(shape, device=dev)
            x0_ptr = x0_buf.data_ptr()

        def step_ops(s):
            if per_sample_time:
                eng.run_time_path(s)
            else:
                _lib.check(lib.ddm_select_row(ss_table.data_ptr(), counter.data_ptr(), eng.ss.data_ptr(), eng.ss_width, s))
            eng.run_body(s)
            if guided is not None:              # dd:710-777: raw eps, optional clamp, guide blended into the known region
                _lib.check(lib.ddm_sampler_step_guided(eng.x.data_ptr(), eng.out.data_ptr(), noise_ptr, noise_stride,
                                                       _p(guided["guide"]), _p(guided["mask"]), _p(guided["guide_noise"]),
                                                       numel if guided["guide_noise"] is not None else 0, x0_ptr,
                                                       coef_dev.data_ptr(), counter.data_ptr(), 1, obj, 1 if guided["clip"] else 0,
                                                       seed, numel, s))
            elif kind == _KIND_DDPM_LEARNED:    # eng.out is [B, 2C, H, W] = (pred_noise | variance fraction)
                _lib.check(lib.ddm_sampler_step_learned(eng.x.data_ptr(), eng.out.data_ptr(), noise_ptr, noise_stride, x0_ptr,
                                                        coef_dev.data_ptr(), counter.data_ptr(), 1, seed, numel, C * H * W, s))
            else:
                _lib.check(lib.ddm_sampler_step(kind, eng.x.data_ptr(), eng.out.data_ptr(), noise_ptr, noise_stride, x0_ptr,
                                                coef_dev.data_ptr(), counter.data_ptr(), 1, obj, seed, numel, s))

        imgs = [x_T] if return_all_timesteps else None
        self._last_graph_launches = 0
        if eager:
            eng.x.copy_(x_T)
            for i, t in enumerate(times):
                if per_sample_time:
                    eng.time.fill_(float(t))
                x_t = eng.x.clone() if trace is not None else None
                step_ops(stream.cuda_stream)
                if trace is not None:
                    trace.append(dict(t=t, x_t=x_t, model_out=eng.out.clone(), x_next=eng.x.clone(),
                                      x_start=(x0_buf if x0_buf is not None else eng.x_self_cond).clone()))
                if imgs is not None:
                    imgs.append(eng.x.clone())
        else:
            if loop["graph"] is None:
                eng.x.copy_(x_T)                                  # warm-up launch outside capture, then restore state
                step_ops(stream.cuda_stream)
                counter.copy_(torch.tensor([0, loop["epoch"]], dtype=torch.int32))
                if eng.x_self_cond is not None:
                    eng.x_self_cond.zero_()
                # The WHOLE step loop is one CUDA graph: `chunk` steps are captured back to back (the device-side step
                # counter makes every step of the capture pick its own table rows), so a sampling call is S / chunk host
                # launches -- one for S <= MAX_GRAPH_STEPS (DDIM-50 / -100 / -250).  Longer loops (the 1000-step ancestral
                # sampler) replay a chunk that divides S, keeping the instantiated graph at a few thousand kernel nodes.
                chunk = S if S <= MAX_GRAPH_STEPS else max(d for d in range(1, MAX_GRAPH_STEPS + 1) if S % d == 0)
                graph = torch.cuda.CUDAGraph()
                torch.cuda.synchronize(dev)
                n0 = _lib.launch_count()
                with torch.cuda.graph(graph):
                    for _ in range(chunk):
                        step_ops(torch.cuda.current_stream(dev).cuda_stream)
                loop["per_step"] = (_lib.launch_count() - n0) // chunk
                loop["graph"], loop["chunk"] = graph, chunk
                loop["keep"] = step_noise                         # raw pointers baked into the graph
            eng.x.copy_(x_T)
            graph = loop["graph"]
            for _ in range(S // loop["chunk"]):
                graph.replay()
            self._last_graph_launches = loop["per_step"] * S      # kernels launched by replays (not via the C-ABI counter)
            self._last_host_launches = S // loop["chunk"]
        ret = eng.x.clone() if imgs is None else torch.stack(imgs, dim=1)
        out = torch.empty_like(ret)
        _lib.check(lib.ddm_finalize(ret.data_ptr(), out.data_ptr(), 1 if (self._auto_normalize and not raw) else 0, ret.numel(),
                                    torch.cuda.current_stream(dev).cuda_stream))
        return out

    # ------------------------------------------------------------------ public sampling API (dd:638-708, 779-783)
    @torch.no_grad()
    def p_sample(self, x, t: int, x_self_cond=None, **cond):
        """One ancestral step on an arbitrary x (dd:638-645); returns (pred_img, x_start)."""
        b = x.shape[0]
        bt = torch.full((b,), t, device=x.device, dtype=torch.long)
        mean, _, logvar, x_start = self.p_mean_variance(x=x, t=bt, x_self_cond=x_self_cond, clip_denoised=True, **cond)
        noise = torch.randn_like(x) if t > 0 else 0.
        return mean + (0.5 * logvar).exp() * noise, x_start

    @torch.no_grad()
    def p_sample_loop(self, shape, return_all_timesteps=False, *, noise=None, step_noise=None, use_graph=True, trace=None):
        """dd:647-664.  Keyword-only extras: `noise` = x_T, `step_noise` = per-step draws (parity mode)."""
        times = list(reversed(range(self.num_timesteps)))
        return self._run_loop(_KIND_DDPM, tuple(shape), times, self._ddpm_coefs(times), x_T=noise, step_noise=step_noise,
                              return_all_timesteps=return_all_timesteps, use_graph=use_graph, trace=trace)

    @torch.no_grad()
    def ddim_sample(self, shape, sampling_timesteps=None, return_all_timesteps=False, *, noise=None, step_noise=None,
                    use_graph=True, trace=None):
        """dd:666-708."""
        S = self.sampling_timesteps if sampling_timesteps is None else sampling_timesteps
        pairs = self._ddim_pairs(S)
        return self._run_loop(_KIND_DDIM, tuple(shape), [t for t, _ in pairs], self._ddim_coefs(pairs, self.ddim_sampling_eta),
                              x_T=noise, step_noise=step_noise, return_all_timesteps=return_all_timesteps,
                              use_graph=use_graph, trace=trace)

    @torch.no_grad()
    def ddim_sample_guided(self, shape, sampling_timesteps=None, guide=None, mask=None, clip_denoised=True, *, noise=None,
                           step_noise=None, guide_noise=None, trace=None):
        """dd:710-777: DDIM that keeps a known region (`mask` == 0) on the guide image: after every non-final update
        `img = img * mask + q_sample(guide, t) * (1 - mask)`.  Like the reference it uses the network's raw noise prediction
        (no re-derivation after the x0 clamp), clamps x0 only if `clip_denoised`, and always unnormalises the result.
        The reference's inline matplotlib display of every step is not reproduced.  Keyword-only extras: `noise` = x_T,
        `step_noise` / `guide_noise` = per-step draws for the DDIM noise / the guide's q_sample (parity mode)."""
        S = self.sampling_timesteps if sampling_timesteps is None else sampling_timesteps
        shape = tuple(shape)
        dev = self.device
        pairs = self._ddim_pairs(S)
        if (guide is None) != (mask is None):
            raise ValueError("guide and mask go together")
        g = dict(guide=None, mask=None, guide_noise=None, clip=bool(clip_denoised))
        if guide is not None:
            g["guide"] = guide.to(dev, torch.float32).expand(shape).contiguous()
            g["mask"] = mask.to(dev, torch.float32).expand(shape).contiguous()
            if guide_noise is not None:
                gn = guide_noise.to(dev, torch.float32).contiguous()
                pad = len(pairs) - gn.shape[0]
                if pad > 0:
                    gn = torch.cat([gn, torch.zeros((pad,) + shape, device=dev)], dim=0)
                g["guide_noise"] = gn
        out = self._run_loop(_KIND_DDIM, shape, [t for t, _ in pairs], self._ddim_coefs(pairs, self.ddim_sampling_eta), x_T=noise,
                             step_noise=step_noise, trace=trace, raw=True, guided=g)
        return unnormalize_to_zero_to_one(out)          # dd:776 (unconditionally, also for auto_normalize=False models)

    @torch.no_grad()
    def q_sample(self, x_start, t, noise=None):
        """dd:805-821 (forward process; elementwise glue outside the sampling loop)."""
        noise = torch.randn_like(x_start) if noise is None else noise
        shape = (x_start.shape[0],) + (1,) * (x_start.ndim - 1)
        a = self.sqrt_alphas_cumprod.gather(-1, t).reshape(shape)
        b = self.sqrt_one_minus_alphas_cumprod.gather(-1, t).reshape(shape)
        return a * x_start + b * noise

    @torch.no_grad()
    def interpolate(self, x1, x2, t=None, lam=0.5, *, q_noise=None, step_noise=None, use_graph=True, cond=None, text_emb=None):
        """dd:785-803: both images noised to step `t`, blended with weight `lam`, then ancestral steps t-1 .. 0 on the
        B200 path.  Like the reference, inputs are taken as already normalised and the result is returned raw.
        Keyword-only extras: `q_noise` = (noise for x1, noise for x2), `step_noise` = per-step draws (parity mode),
        `cond` / `text_emb` = the condition of a conditional model (the subclasses put it in the reference's position)."""
        assert x1.shape == x2.shape
        t = self.num_timesteps - 1 if t is None else int(t)
        tb = torch.full((x1.shape[0],), t, device=x1.device, dtype=torch.long)
        n1, n2 = (None, None) if q_noise is None else q_noise
        img = (1 - lam) * self.q_sample(x1, tb, n1) + lam * self.q_sample(x2, tb, n2)
        times = list(reversed(range(0, t)))
        return self._run_loop(_KIND_DDPM, tuple(img.shape), times, self._ddpm_coefs(times), x_T=img.float().contiguous(),
                              step_noise=step_noise, use_graph=use_graph, raw=True, cond=cond, text_emb=text_emb)

    @torch.no_grad()
    def sample(self, batch_size=16, return_all_timesteps=False, **kw):
        """dd:779-783."""
        (h, w), channels = self.image_size, self.channels
        fn = self.p_sample_loop if not self.is_ddim_sampling else self.ddim_sample
        return fn((batch_size, channels, h, w), return_all_timesteps=return_all_timesteps, **kw)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("training (p_losses, dd:823-900) is outside the B200 sampling hot path; "
                                  "train with the reference and load its state_dict here")


_KIND_DDIM, _KIND_DDPM, _KIND_DDPM_LEARNED = 0, 1, 2

GaussianDiffusion = DenoisingDiffusion      # upstream lucidrains name
