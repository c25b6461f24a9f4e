"""Multi-GPU path on CPU: world_size-2 gloo run of the batch-sharding logic (no collective on the data
path; the gather here is only the checker)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ir2rgb_b200.sharding import shard, shard_bounds


def test_shard_bounds_tile_the_batch():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _per_pair_work(frames):
    """The hot path's operators on a batch chunk, through the CPU oracle (there is no CPU product path): the warp /
    brightness-error stage (Resample2d -> subtract -> ChannelNorm, models.py:109-111) and a small Correlation.  Every
    operator treats batch rows independently, so the sharded result must equal the unsharded one bit for bit."""
    from oracle import torch_ref as tr
    flow = 2 * torch.tanh(frames[:, :2].flip(2))
    warped = tr.resample2d(frames.contiguous(), flow.contiguous())
    err = tr.channelnorm(frames - warped)
    corr = tr.correlation(frames, warped, 2, 1, 2, 1, 2)
    return torch.cat((err.flatten(1), corr.flatten(1)), 1).sum(1)


def _worker(rank, world, port, n_items, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                   # every rank sees the same synthetic batch
    frames = torch.randn(n_items, 3, 8, 8)
    mine = shard(frames, rank, world)
    local = _per_pair_work(mine)
    sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.numel()]))
    parts = [torch.zeros(int(s.item())) for s in sizes]
    if rank == 0:
        dist.gather(local, parts, dst=0) if len(set(int(s) for s in sizes)) == 1 else None
    elif len(set(int(s) for s in sizes)) == 1:
        dist.gather(local, dst=0)
    dist.barrier()
    if rank == 0:
        ok_sizes = sum(int(s) for s in sizes) == n_items
        full = _per_pair_work(frames)
        same = True
        if len(set(int(s) for s in sizes)) == 1:
            same = torch.equal(torch.cat(parts), full)
        ret["ok"] = bool(ok_sizes and same)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [64, 8])
def test_two_rank_gloo_sharding_matches_single_rank(n_items):
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_items, ret), nprocs=world, join=True)
        assert ret.get("ok") is True
