"""Trainer-side integration: let the reference's own `Trainer` sample on the B200 path.

`Trainer.train()` (denoising_diffusion.py:1192-1219) calls `self.ema.ema_model.sample(batch_size=n)` on the averaged copy
of the very object it trains -- grids every `save_and_sample_every` steps and `num_fid_samples` images for FID.  That
object is the reference's (training-capable) `DenoisingDiffusion`; this package's class is sampling-only.  The adapter
keeps both: training stays with the reference, and `sample()` / `ddim_sample()` / `p_sample_loop()` of the EMA copy are
re-pointed at a fast sampler that mirrors the EMA weights -- one `load_state_dict` (weight re-pack) whenever the source's
tensors changed, which is detected from their in-place version counters (EMA updates are in-place `lerp_`/`copy_`).

    trainer = Trainer(diffusion, folder, ...)                      # the reference, unchanged
    fast = diffusion_models_b200.DenoisingDiffusion(diffusion_models_b200.Unet(dim=64, ...), image_size=32, ...).cuda()
    attach_fast_sampler(trainer.ema.ema_model, fast)               # milestone sampling now runs on the B200 kernels
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch


def _fingerprint(module: torch.nn.Module) -> Tuple:
    """Changes whenever any parameter / buffer of `module` is written in place or replaced."""
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


class FastSamplerBinding:
    def __init__(self, source: torch.nn.Module, fast: torch.nn.Module, version_fn: Optional[Callable[[], object]] = None):
        self.source, self.fast = source, fast
        self.version_fn = version_fn or (lambda: _fingerprint(source))
        self._loaded = None
        self.syncs = 0

    def sync(self) -> bool:
        """Mirror the source's weights into the fast sampler if they changed since the last call."""
        v = self.version_fn()
        if v == self._loaded:
            return False
        sd = {k: t.detach() for k, t in self.source.state_dict().items()}
        missing, unexpected = self.fast.load_state_dict(sd, strict=False)
        # the schedule buffers and every network tensor must be there; extras of the source (e.g. training-only buffers) may not
        missing = [k for k in missing if k != "loss_weight"]
        if missing:
            raise KeyError(f"source module lacks tensors the sampler needs: {missing[:5]}{' ...' if len(missing) > 5 else ''}")
        self._loaded = v
        self.syncs += 1
        return True

    def __getattr__(self, name):            # forwards sample / ddim_sample / p_sample_loop / interpolate ... after a sync
        fn = getattr(self.fast, name)
        if not callable(fn):
            return fn

        def call(*args, **kwargs):
            self.sync()
            return fn(*args, **kwargs)
        return call


def attach_fast_sampler(source: torch.nn.Module, fast: torch.nn.Module, version_fn: Optional[Callable[[], object]] = None,
                        methods=("sample", "ddim_sample", "p_sample_loop")) -> FastSamplerBinding:
    """Re-point the sampling methods of `source` (the reference's EMA `DenoisingDiffusion`, used by its `Trainer`,
    `FIDEvaluation` and the sampling scripts) at `fast`, keeping the weights in sync.  Returns the binding (its `.syncs`
    counts the re-packs).  `source` keeps training / `state_dict()` / `forward()` exactly as before."""
    binding = FastSamplerBinding(source, fast, version_fn)
    for m in methods:
        if hasattr(fast, m):
            object.__setattr__(source, m, getattr(binding, m))
    return binding
