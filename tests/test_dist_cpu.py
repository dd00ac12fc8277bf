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


# ------------------------------------------------------------------ training: one all-reduce over the flat gradient buffer
def _grad_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fullsubnet_b200.fullsubnet.model import Model
    torch.manual_seed(0)  # identical replicas
    m = Model(num_freqs=9, look_ahead=1, sequence_model="LSTM", fb_num_neighbors=0, sb_num_neighbors=2,
              fb_output_activate_function="ReLU", sb_output_activate_function=False, fb_model_hidden_size=8,
              sb_model_hidden_size=4, weight_init=False)
    for i, p in enumerate(m.parameters()):  # rank-dependent gradients, as after a local backward
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    flat = m.flat_grad()
    assert flat.numel() == sum(p.numel() for p in m.parameters())
    assert all(p.grad.data_ptr() == flat.data_ptr() + 4 * off for p, off in zip(m.parameters(), m._flat_offsets))
    dist.all_reduce(flat)  # the single collective of the step; the 1/world mean is folded into the optimiser
    for i, p in enumerate(m.parameters()):
        assert torch.equal(p.grad, torch.full_like(p, 3.0 * (i + 1)))  # (1 + 2) * (i + 1): views saw the reduction
    assert m.flat_grad().data_ptr() == flat.data_ptr()  # second call keeps the buffer
    m.zero_grad(set_to_none=False)
    assert float(flat.abs().sum()) == 0.0
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_flat_gradient_allreduce():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_grad_worker, args=(2, port), nprocs=2, join=True)


# ------------------------------------------------------------------ trainer start-up: every rank starts from rank 0's weights
def _bcast_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fullsubnet_b200.fullsubnet.model import Model
    from fullsubnet_b200.trainer import broadcast_parameters, unwrap
    torch.manual_seed(100 + rank)  # DIFFERENT initial weights per rank: without the broadcast the replicas diverge
    m = Model(num_freqs=9, look_ahead=1, sequence_model="LSTM", fb_num_neighbors=0, sb_num_neighbors=2,
              fb_output_activate_function="ReLU", sb_output_activate_function=False, fb_model_hidden_size=8,
              sb_model_hidden_size=4, weight_init=True)
    before = [p.detach().clone() for p in m.parameters()]
    broadcast_parameters(m, dist, src=0)
    gathered = [None] * world
    dist.all_gather_object(gathered, [p.detach().clone() for p in m.parameters()])
    for a, b in zip(gathered[0], gathered[1]):
        assert torch.equal(a, b)
    if rank == 0:
        assert all(torch.equal(a, b) for a, b in zip(before, m.parameters()))  # rank 0 keeps its own
    else:
        assert any(not torch.equal(a, b) for a, b in zip(before, m.parameters()))
    ddp = torch.nn.parallel.DistributedDataParallel(torch.nn.Linear(2, 2))
    assert unwrap(ddp) is ddp.module and unwrap(m) is m
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_parameter_broadcast():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_bcast_worker, args=(2, port), nprocs=2, join=True)
