"""world_size-2 gloo test of the clip-sharding host logic (no GPU): shards cover the batch exactly once,
results gathered on rank 0 are in clip order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fullsubnet_b200.dist import enhance_sharded, gather_waves, shard_bounds


def test_shard_bounds_cover_batch():
    for n in (0, 1, 2, 5, 8, 257):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    noisy = torch.randn(n, 64)
    fake_enhance = lambda x: 2.0 * x + 1.0  # stands in for Inferencer.enhance_batch (elementwise per clip)
    local = enhance_sharded(fake_enhance, noisy, world, rank)
    out = gather_waves(local, n)
    if rank == 0:
        assert torch.equal(out, fake_enhance(noisy))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_enhance():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, 5), nprocs=2, join=True)
