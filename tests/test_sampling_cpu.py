"""Caller integration (diffusion-models_b200/sampling.py) on CPU: batching rule, milestone discovery, ingestion of a
checkpoint in the reference's `Trainer.save` layout (denoising_diffusion.py:1100-1113, `ema` = ema_pytorch state_dict),
and the multi-rank group deal + gather with a stand-in sampler (2-rank gloo)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffusion_models_b200 as ddm          # noqa: E402
from diffusion_models_b200 import sampling   # noqa: E402


def test_num_to_groups_matches_reference_rule():
    # utils.py:30-36: full groups then the remainder
    assert sampling.num_to_groups(25, 8) == [8, 8, 8, 1]
    assert sampling.num_to_groups(16, 16) == [16]
    assert sampling.num_to_groups(5, 64) == [5]
    assert sampling.num_to_groups(0, 4) == []
    assert sum(sampling.num_to_groups(1000, 64)) == 1000


def test_find_milestones(tmp_path):
    for name in ("model-10.pt", "model-2.pt", "model-x.pt", "sample-3.png", "model-7.pt.bak"):
        (tmp_path / name).write_bytes(b"")
    assert sampling.find_milestones(tmp_path) == [2, 10]


def _tiny_diffusion(seed):
    torch.manual_seed(seed)
    model = ddm.Unet(dim=16, dim_mults=(1, 2), channels=3)
    for p in model.parameters():
        torch.nn.init.normal_(p, std=0.05)
    return ddm.DenoisingDiffusion(model, image_size=16, timesteps=20, sampling_timesteps=4)


def _trainer_checkpoint(diffusion_ema, diffusion_online, step=1234):
    """What Trainer.save writes: ema_pytorch.EMA.state_dict() has online_model.*, ema_model.*, initted, step."""
    ema = {"initted": torch.tensor(True), "step": torch.tensor(step)}
    ema.update({"online_model." + k: v.clone() for k, v in diffusion_online.state_dict().items()})
    ema.update({"ema_model." + k: v.clone() for k, v in diffusion_ema.state_dict().items()})
    return {"step": step, "model": {k: v.clone() for k, v in diffusion_online.state_dict().items()}, "opt": {}, "ema": ema,
            "scaler": None, "version": "2.0.0"}


def test_checkpoint_ingestion_ema_and_raw(tmp_path):
    a, b, target = _tiny_diffusion(1), _tiny_diffusion(2), _tiny_diffusion(3)
    ckpt = _trainer_checkpoint(a, b)
    path = tmp_path / "model-5.pt"
    torch.save(ckpt, path)
    assert sampling.load_checkpoint(target, path) == 1234                 # EMA copy by default (sampling.py:157-159)
    for k, v in a.state_dict().items():
        assert torch.equal(target.state_dict()[k], v), k
    assert sampling.load_checkpoint(target, ckpt, use_ema=False) == 1234   # raw training copy
    for k, v in b.state_dict().items():
        assert torch.equal(target.state_dict()[k], v), k
    sampling.load_checkpoint(target, a.state_dict())                        # a bare state_dict passes through
    with pytest.raises(KeyError):
        sampling.extract_state_dict({"ema": {"initted": torch.tensor(True)}, "step": 1})
    with pytest.raises(RuntimeError):                                       # strict: the reference's key set exactly
        bad = dict(a.state_dict())
        bad.pop(next(iter(bad)))
        sampling.load_checkpoint(target, bad)


class _FakeDiffusion(torch.nn.Module):
    """Stand-in sampler: every image is filled with a running counter so that order and counts can be checked."""

    def __init__(self, rank=0):
        super().__init__()
        self.image_size, self.channels, self.rank = (4, 4), 3, rank
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.calls = []

    def sample(self, batch_size=16):
        self.calls.append(("sample", batch_size))
        return torch.full((batch_size, 3, 4, 4), float(self.rank * 1000 + len(self.calls)))

    def ddim_sample(self, shape, sampling_timesteps=None):
        self.calls.append(("ddim", shape[0], sampling_timesteps))
        return torch.full(tuple(shape), float(self.rank * 1000 + len(self.calls)))


def test_generate_samples_single_process():
    d = _FakeDiffusion()
    out = sampling.generate_samples(d, 25, 8)
    assert out.shape == (25, 3, 4, 4) and d.calls == [("sample", 8)] * 3 + [("sample", 1)]
    assert out[:8].unique().tolist() == [1.0] and out[24].unique().tolist() == [4.0]
    d = _FakeDiffusion()
    sampling.generate_samples(d, 10, 4, ddim_sampling_timesteps=7)
    assert d.calls == [("ddim", 4, 7), ("ddim", 4, 7), ("ddim", 2, 7)]


def test_assign_groups_round_robin():
    assert sampling.assign_groups([8, 8, 8, 1], 2) == [[0, 2], [1, 3]]
    assert sampling.assign_groups([5], 4) == [[0], [], [], []]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffusion_models_b200 import sampling as smp
    d = _FakeDiffusion(rank)
    out = smp.generate_samples(d, 21, 8)              # groups [8, 8, 5]: rank 0 -> groups 0 and 2, rank 1 -> group 1
    torch.save((out, d.calls), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_generate_samples_two_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(2)]
    assert outs[0][1] == [("sample", 8), ("sample", 5)] and outs[1][1] == [("sample", 8)]
    for out, _ in outs:                                # every rank holds the pool in group order
        assert out.shape == (21, 3, 4, 4)
        assert out[:8].unique().tolist() == [1.0]      # rank 0, first call
        assert out[8:16].unique().tolist() == [1001.0] # rank 1, first call
        assert out[16:].unique().tolist() == [2.0]     # rank 0, second call


def test_trainer_adapter_syncs_only_when_weights_change():
    """attach_fast_sampler (trainer_adapter.py): the EMA copy's sample() is re-pointed at the fast sampler and the weights are
    mirrored once per change (version counters), not once per call -- dd:1192-1219 calls sample() many times per milestone."""
    import diffusion_models_b200 as ddm
    src = ddm.DenoisingDiffusion(ddm.Unet(dim=16, dim_mults=(1, 2)), image_size=16, sampling_timesteps=2)     # stands in for the
    fast = ddm.DenoisingDiffusion(ddm.Unet(dim=16, dim_mults=(1, 2)), image_size=16, sampling_timesteps=2)    # reference's EMA copy
    calls = []
    fast.sample = lambda batch_size=16, return_all_timesteps=False: calls.append(batch_size) or torch.zeros(batch_size, 3, 16, 16)
    b = ddm.attach_fast_sampler(src, fast)
    assert src.sample(batch_size=4).shape == (4, 3, 16, 16) and src.sample(batch_size=2).shape == (2, 3, 16, 16)
    assert calls == [4, 2] and b.syncs == 1
    w = "model.init_conv.weight"
    assert torch.equal(fast.state_dict()[w], src.state_dict()[w])
    with torch.no_grad():
        for p in src.parameters():
            p.lerp_(torch.zeros_like(p), 0.5)                     # an EMA-style in-place update
    src.sample(batch_size=1)
    assert b.syncs == 2 and torch.equal(fast.state_dict()[w], src.state_dict()[w])
    src.sample(batch_size=1)
    assert b.syncs == 2
