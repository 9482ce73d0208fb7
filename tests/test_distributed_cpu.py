"""N > 1 host logic on CPU: contiguous batch sharding + the single all-gather, with a 2-rank gloo group.
The per-rank sampler is a stand-in function (the CUDA sampler needs a B200); what is checked is that every rank ends
up with the full batch in global row order, including ragged batches, and that per-rank seeds differ."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import diffusion_models_b200 as ddm
    bounds = ddm.shard_bounds(batch, world)

    def sampler(local_batch, r):
        lo, hi = bounds[r]
        if batch < world:           # fewer samples than ranks: every rank draws one row, the first `batch` are kept
            lo, hi = r, r + 1
        assert hi - lo == local_batch
        rows = torch.arange(lo, hi, dtype=torch.float32)
        return rows[:, None, None, None].expand(-1, 3, 4, 4).contiguous() + 0.5

    full = ddm.sample_sharded(sampler, batch)
    # the usual torchrun pattern: every rank seeds identically -- the per-call salt (x_T seed + step-noise salt) must differ
    torch.manual_seed(0)
    salt_a = ddm.DenoisingDiffusion._call_salt()
    salt_b = ddm.DenoisingDiffusion._call_salt()
    torch.manual_seed(0)
    assert ddm.DenoisingDiffusion._call_salt() == salt_a and salt_a != salt_b      # reproducible after re-seeding
    torch.save(dict(full=full, salt=salt_a), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7, 1])
def test_sample_sharded_two_ranks(tmp_path, batch):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    want = (torch.arange(batch, dtype=torch.float32) + 0.5)[:, None, None, None].expand(-1, 3, 4, 4)
    salts = []
    for r in range(2):
        rec = torch.load(os.path.join(tmp_path, f"rank{r}.pt"))
        got = rec["full"]
        assert got.shape == (batch, 3, 4, 4)
        assert torch.equal(got, want)
        salts.append(rec["salt"])
    assert salts[0] != salts[1], "ranks seeded identically must still draw different x_T / step noise"


def test_shard_bounds_cover_batch():
    import diffusion_models_b200 as ddm
    for batch in (0, 1, 5, 16, 1024, 1023):
        for world in (1, 2, 4, 8):
            b = ddm.shard_bounds(batch, world)
            assert b[0][0] == 0 and b[-1][1] == batch
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_passthrough():
    import diffusion_models_b200 as ddm
    out = ddm.sample_sharded(lambda b, r: torch.full((b, 2), float(r)), 5)
    assert out.shape == (5, 2)
