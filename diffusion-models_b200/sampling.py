"""Caller integration: what the reference's `sampling.py` scripts and `Trainer`'s periodic
sampling do around the hot path, driven by the B200 sampler.

Behaviour mirrored (not code): `denoising-diffusion-pytorch/sampling.py:120-194` (milestone discovery `model-(\\d+).pt`,
`data["ema"]` ingestion, `num_to_groups` batching, a `sample-{milestone}.png` grid, a pool of samples for FID/IS),
`denoising_diffusion/utils.py:30-36` (`num_to_groups`), `denoising_diffusion.py:1100-1113` (`Trainer.save` layout:
`{step, model, opt, ema, scaler, version}`) and `:1198-1219` (periodic sampling inside `Trainer.train`).

Not rebuilt: FID / Inception-Score scoring (pytorch_fid / torchvision Inception weights are not available and are
outside the path, SURVEY section 8 "out of scope"); the sample pool is written to `fid_samples-{milestone}.npz` so the
reference's scorers can be pointed at it.

New relative to the reference: the sample pool is generated on all GPUs of the node (`torchrun`), groups dealt
round-robin to the ranks, one gather at the end (the reference samples on the main process only).
"""
from __future__ import annotations

import argparse
import math
import os
import re
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Union

import torch
import torch.distributed as dist

_MILESTONE = re.compile(r"model-(\d+)\.pt")


def num_to_groups(num: int, divisor: int) -> List[int]:
    """`num` split into groups of `divisor` plus a remainder group (utils.py:30-36): 25, 8 -> [8, 8, 8, 1]."""
    if divisor <= 0:
        raise ValueError("divisor must be positive")
    groups, remainder = divmod(int(num), int(divisor))
    return [divisor] * groups + ([remainder] if remainder > 0 else [])


def find_milestones(folder: Union[str, os.PathLike]) -> List[int]:
    """Sorted milestone numbers of the `model-N.pt` files in a results folder (sampling.py:120-129)."""
    out = []
    for name in os.listdir(folder):
        m = _MILESTONE.fullmatch(name)
        if m:
            out.append(int(m.group(1)))
    return sorted(out)


def _strip_prefix(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def extract_state_dict(data: Dict, use_ema: bool = True) -> Dict[str, torch.Tensor]:
    """The diffusion module's `state_dict` out of a `Trainer.save` checkpoint.

    `data["ema"]` is `ema_pytorch.EMA.state_dict()`: keys `ema_model.<k>` (the averaged copy the reference samples
    from, sampling.py:157-159), `online_model.<k>`, plus the scalars `initted` / `step`.  `data["model"]` is the raw
    training copy.  A bare `state_dict` (no `model`/`ema` entries) is passed through.
    """
    if not isinstance(data, dict):
        raise TypeError("checkpoint must be a dict")
    if use_ema and isinstance(data.get("ema"), dict):
        sd = _strip_prefix(data["ema"], "ema_model.")
        if not sd:
            raise KeyError("checkpoint['ema'] holds no 'ema_model.*' entries")
        return sd
    if isinstance(data.get("model"), dict):
        return dict(data["model"])
    if "ema" in data or "model" in data:
        raise KeyError("checkpoint has no usable 'ema' / 'model' state_dict")
    return dict(data)


def load_checkpoint(diffusion: torch.nn.Module, checkpoint: Union[str, os.PathLike, Dict], use_ema: bool = True,
                    strict: bool = True) -> int:
    """Load a reference checkpoint (`model-N.pt` path or the already-loaded dict) into a `DenoisingDiffusion`-family
    module of this package.  Returns the training step stored in the checkpoint (0 if absent).  Key names and shapes
    are the reference's, so this is a plain `load_state_dict`; the packed bf16 weights are rebuilt lazily."""
    if not isinstance(checkpoint, dict):
        checkpoint = torch.load(str(checkpoint), map_location="cpu", weights_only=True)
    sd = extract_state_dict(checkpoint, use_ema=use_ema)
    diffusion.load_state_dict(sd, strict=strict)
    step = checkpoint.get("step", 0) if isinstance(checkpoint, dict) else 0
    return int(step) if not torch.is_tensor(step) else int(step.item())


def _dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def assign_groups(groups: List[int], world: int) -> List[List[int]]:
    """Round-robin deal of sample groups to ranks: rank r gets groups r, r+world, ...  (indices into `groups`)."""
    return [list(range(r, len(groups), world)) for r in range(world)]


@torch.no_grad()
def generate_samples(diffusion, num_samples: int, batch_size: int, ddim_sampling_timesteps: Optional[int] = None,
                     sample_kwargs: Optional[dict] = None) -> torch.Tensor:
    """`num_samples` images in groups of `batch_size` (sampling.py:163-170, 176-189).  `ddim_sampling_timesteps=None`
    follows the module's own `sample()` dispatch; an int forces `ddim_sample(..., sampling_timesteps=S)`.

    With an initialised process group the groups are dealt round-robin to the ranks and gathered once at the end;
    every rank returns the full `[num_samples, C, H, W]` tensor in group order."""
    kw = dict(sample_kwargs or {})
    groups = num_to_groups(num_samples, batch_size)
    world, rank = _dist_info()
    mine = assign_groups(groups, world)[rank]
    size = diffusion.image_size
    h, w = (size, size) if isinstance(size, int) else tuple(size)
    c = diffusion.channels

    def one(n: int) -> torch.Tensor:
        if ddim_sampling_timesteps is not None:
            return diffusion.ddim_sample((n, c, h, w), sampling_timesteps=ddim_sampling_timesteps, **kw)
        return diffusion.sample(batch_size=n, **kw)

    local = [one(groups[i]) for i in mine]
    if world == 1:
        return torch.cat(local, dim=0)
    # one collective: every rank contributes a [max_rows, C, H, W] block (ragged tails zero-padded)
    dev = local[0].device if local else next(diffusion.parameters()).device
    rows = [sum(groups[i] for i in assign_groups(groups, world)[r]) for r in range(world)]
    block = torch.zeros((max(rows), c, h, w), dtype=torch.float32, device=dev)
    if local:
        block[: rows[rank]] = torch.cat(local, dim=0)
    out = torch.empty((world * max(rows), c, h, w), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(out, block)
    per_rank = [out[r * max(rows): r * max(rows) + rows[r]] for r in range(world)]
    # back to group order
    offs = [0] * world
    parts = []
    for i, n in enumerate(groups):
        r = i % world
        parts.append(per_rank[r][offs[r]: offs[r] + n])
        offs[r] += n
    return torch.cat(parts, dim=0)


def save_image_grid(images: torch.Tensor, path: Union[str, os.PathLike], nrow: Optional[int] = None) -> None:
    """`torchvision.utils.save_image` grid like sampling.py:172 (nrow = floor(sqrt(N)) by default)."""
    from torchvision import utils as tv_utils

    n = images.shape[0]
    tv_utils.save_image(images.detach().float().cpu(), str(path), nrow=nrow or max(1, int(math.sqrt(n))))


def run(diffusion, trained_models_folder: Union[str, os.PathLike], out_folder: Union[str, os.PathLike],
        milestones: Optional[Iterable[int]] = None, num_samples: int = 25, batch_size: int = 64,
        ddim_sampling_timesteps: Optional[int] = None, num_fid_samples: int = 0, use_ema: bool = True) -> List[dict]:
    """For each milestone: load `model-{m}.pt`, write `sample-{m}.png` and (optionally) `fid_samples-{m}.npz`."""
    out_folder = Path(out_folder)
    world, rank = _dist_info()
    if rank == 0:
        out_folder.mkdir(parents=True, exist_ok=True)
    ms = list(milestones) if milestones is not None else find_milestones(trained_models_folder)
    report = []
    for m in ms:
        step = load_checkpoint(diffusion, Path(trained_models_folder) / f"model-{m}.pt", use_ema=use_ema)
        diffusion.eval()
        imgs = generate_samples(diffusion, num_samples, batch_size, ddim_sampling_timesteps)
        entry = {"milestone": m, "step": step, "samples": tuple(imgs.shape)}
        if rank == 0:
            save_image_grid(imgs, out_folder / f"sample-{m}.png")
        if num_fid_samples > 0:
            pool = generate_samples(diffusion, num_fid_samples, batch_size, ddim_sampling_timesteps)
            if rank == 0:
                import numpy as np

                np.savez_compressed(out_folder / f"fid_samples-{m}.npz", images=pool.cpu().numpy())
            entry["fid_samples"] = tuple(pool.shape)
        report.append(entry)
    return report


def main(argv=None):
    """CLI with the reference script's flags (sampling.py:46-109) plus the model hyper-parameters it hard-codes."""
    from . import Unet, DenoisingDiffusion

    ap = argparse.ArgumentParser(description="Sample from trained checkpoints with the B200 sampler")
    ap.add_argument("--trained_models_folder", type=str, default="./results")
    ap.add_argument("--model", type=int, default=None, help="milestone number; default: every model-N.pt in the folder")
    ap.add_argument("--generation_results_folder", type=str, default=None)
    ap.add_argument("--ddim_sampling_timesteps", type=int, default=None, help="default: ancestral sampling, as in the reference script")
    ap.add_argument("--num_samples", type=int, default=25)
    ap.add_argument("--batch_size", type=int, default=64)
    ap.add_argument("--num_fid_samples", type=int, default=1000)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--dim_mults", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--image_size", type=int, default=32)
    ap.add_argument("--timesteps", type=int, default=1000)
    ap.add_argument("--raw_model", action="store_true", help="use data['model'] instead of the EMA copy")
    args = ap.parse_args(argv)

    if "RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    model = Unet(dim=args.dim, dim_mults=tuple(args.dim_mults), dropout=0.1)
    diffusion = DenoisingDiffusion(model, image_size=args.image_size, timesteps=args.timesteps).to(dev)
    out = args.generation_results_folder
    if out is None:
        tag = args.ddim_sampling_timesteps if args.ddim_sampling_timesteps is not None else "ddpm"
        out = Path("./results_ddim") / f"{os.path.basename(os.path.normpath(args.trained_models_folder))}_{tag}"
    rep = run(diffusion, args.trained_models_folder, out, [args.model] if args.model is not None else None, args.num_samples,
              args.batch_size, args.ddim_sampling_timesteps, args.num_fid_samples, use_ema=not args.raw_model)
    if _dist_info()[1] == 0:
        for e in rep:
            print(e)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
