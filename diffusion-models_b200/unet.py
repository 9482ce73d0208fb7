"""`Unet` -- drop-in for the reference denoiser (denoising_diffusion.py:233-390) whose forward runs on the
sm_100a kernels of libddm_b200.so.

The module owns fp32 master parameters under exactly the reference's state_dict names (so `load_state_dict` of a
reference checkpoint works unchanged) and a cache of `UnetEngine` plans keyed by input shape.
The forward pass is inference-only (the reference's sampling path runs under `torch.inference_mode`) and CUDA-only:
there is no PyTorch fallback.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from .arch import UnetSpec, build_spec
from .engine import UnetEngine


def _init_param(shape, kind: str) -> torch.Tensor:
    """Same distributions as the reference's nn.Conv2d / nn.Linear / RMSNorm / mem_kv defaults."""
    if kind == "ones":
        return torch.ones(shape)
    if kind == "randn":
        return torch.randn(shape)
    if kind == "conv":
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        bound = 1.0 / math.sqrt(fan_in)          # kaiming_uniform_(a=sqrt(5))
        return torch.empty(shape).uniform_(-bound, bound)
    if kind.startswith("bias:"):
        bound = 1.0 / math.sqrt(int(kind.split(":")[1]))
        return torch.empty(shape).uniform_(-bound, bound)
    raise ValueError(kind)


class _Holder(nn.Module):
    """Parameter container; gives dotted reference names like `downs.0.0.block1.proj.weight` a module tree."""


def _register(root: nn.Module, dotted: str, param: nn.Parameter) -> None:
    *path, leaf = dotted.split(".")
    mod = root
    for part in path:
        nxt = mod._modules.get(part)
        if nxt is None:
            nxt = _Holder()
            mod.add_module(part, nxt)
        mod = nxt
    mod.register_parameter(leaf, param)


class Unet(nn.Module):
    """Same constructor surface as the reference `Unet` (dd:234-252)."""

    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=3, self_condition=False,
                 learned_variance=False, learned_sinusoidal_cond=False, random_fourier_features=False,
                 learned_sinusoidal_dim=16, sinusoidal_pos_emb_theta=10000, dropout=0., attn_dim_head=32, attn_heads=4,
                 full_attn=None, flash_attn=False, **_spec_extra):
        super().__init__()
        if learned_sinusoidal_cond or random_fourier_features:
            # the reference's own DenoisingDiffusion asserts these off (dd:457); they are not on the sampling path
            raise NotImplementedError("learned / random sinusoidal time embeddings are outside the sampling hot path")
        self.channels = channels
        self.self_condition = self_condition
        self.random_or_learned_sinusoidal_cond = False
        self.dropout = dropout            # identity in eval; kept for signature parity
        self.flash_attn = flash_attn      # both settings compute softmax(qk^T)v; one fused kernel serves both
        self.spec: UnetSpec = build_spec(dim, init_dim, out_dim, dim_mults, channels, self_condition, learned_variance,
                                         sinusoidal_pos_emb_theta, attn_dim_head, attn_heads, full_attn, **_spec_extra)
        self.out_dim = self.spec.out_dim
        for name, (shape, kind) in self.spec.params.items():
            _register(self, name, nn.Parameter(_init_param(shape, kind)))
        self._engines: Dict[Tuple, UnetEngine] = {}
        self._stamp = None

    @property
    def downsample_factor(self) -> int:
        return self.spec.downsample_factor

    # ------------------------------------------------------------------ engine cache
    def _weights_stamp(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, batch: int, height: int, width: int, *, time_rows: Optional[int] = None, text_tokens: int = 0,
               device: Optional[torch.device] = None) -> UnetEngine:
        """Plan (packed weights + static activation arena) for this input shape; rebuilt when weights change."""
        p0 = next(self.parameters())
        device = torch.device(device) if device is not None else p0.device
        if device.type != "cuda":
            raise RuntimeError("diffusion_models_b200.Unet runs only on a CUDA (B200) device; move the module with "
                               ".to('cuda') -- there is no CPU fallback")
        stamp = self._weights_stamp()
        if stamp != self._stamp:
            self._engines.clear()
            self._stamp = stamp
        time_rows = batch if time_rows is None else time_rows
        key = (batch, height, width, time_rows, text_tokens, device.index)
        eng = self._engines.get(key)
        if eng is None:
            with torch.no_grad():
                eng = UnetEngine(self.spec, dict(self.named_parameters()), batch, height, width, device,
                                 time_rows=time_rows, text_tokens=text_tokens)
            self._engines[key] = eng
        return eng

    def drop_engines(self) -> None:
        self._engines.clear()

    def __deepcopy__(self, memo):
        # EMA wrappers deep-copy the model (dd:1024); plans hold raw device pointers and are rebuilt lazily instead
        engines, self._engines = self._engines, {}
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            import copy
            for k, v in self.__dict__.items():
                new.__dict__[k] = copy.deepcopy(v, memo)
        finally:
            self._engines = engines
        new._stamp = None
        return new

    # ------------------------------------------------------------------ forward
    def _stage_inputs(self, eng: UnetEngine, x, time, x_self_cond=None, cond=None, text_emb=None):
        eng.x.copy_(x)
        eng.time.copy_(time.reshape(-1)[: eng.time_rows] if time.numel() >= eng.time_rows else time.expand(eng.time_rows))
        if eng.x_self_cond is not None:
            if x_self_cond is None:
                eng.x_self_cond.zero_()               # dd:353 default zeros_like(x)
            else:
                eng.x_self_cond.copy_(x_self_cond)
        if eng.cond is not None:
            if cond is None:
                raise ValueError("this Unet was built with cond_channels > 0: pass cond=")
            assert cond.shape[0] == x.shape[0], "batch mismatch between x and cond"
            eng.cond.copy_(cond)
        if eng.text is not None and text_emb is not None:
            eng.text.copy_(text_emb.reshape(eng.text.shape))

    @torch.no_grad()
    def forward(self, x: torch.Tensor, time: torch.Tensor, x_self_cond: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B,C,H,W) fp32, (B,) int64 -> (B,out_dim,H,W) fp32, dd:349-390."""
        b, _, h, w = x.shape
        assert all(d % self.downsample_factor == 0 for d in (h, w)), \
            f"your input dimensions {(h, w)} need to be divisible by {self.downsample_factor}, given the unet"
        eng = self.engine(b, h, w, device=x.device)
        self._stage_inputs(eng, x, time, x_self_cond)
        eng.run_time_path()
        eng.run_body()
        return eng.out.clone()
