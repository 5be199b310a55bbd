"""CPU test of the data-parallel host logic with gloo, world size 2: batch sharding, the flat gradient
buffer, and the single all-reduce whose 1/world scaling reproduces DDP's gradient average."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(7, 5)
        self.b = torch.nn.Parameter(torch.zeros(3))
        self.shared = self.a            # aliased sub-module, like user_encoder.news_encoder

    def forward(self, x):
        return self.a(x).sum() + (self.b * x[:, :3].mean(0)).sum()


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from nnr_b200.trainer import TrainStep, shard_batch
    torch.manual_seed(0)
    model = _Toy()
    ts = TrainStep(model, world_size=world)
    assert ts.flat.numel() % 4 == 0
    # params and grads are views of the flat buffers
    for p in model.parameters():
        assert ts.flat.data_ptr() <= p.data_ptr() < ts.flat.data_ptr() + ts.flat.numel() * 4
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 7, generator=g)
    batch = shard_batch({'x': x, 'none': None}, rank, world)
    assert batch['none'] is None and batch['x'].shape[0] == 4
    assert torch.equal(batch['x'], x[rank * 4:(rank + 1) * 4])
    ts.zero_grad()
    model(batch['x']).backward()
    local = ts.gflat.clone()
    ts.reduce_gradients()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(ts.gflat, sum(gathered))
    # mean over ranks (grad_scale = 1/world) == gradient of the mean loss over shards
    ref = _Toy()
    ref.load_state_dict(model.state_dict())
    sum(ref(x[r * 4:(r + 1) * 4]) for r in range(world)).div(world).backward()
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p.grad / world, q.grad, atol=1e-6), n
    try:
        ts.optimizer_step()
        raised = False
    except RuntimeError:
        raised = True
    assert raised, 'optimizer_step must refuse to run without CUDA'
    out.put(rank)
    dist.destroy_process_group()


def test_dp_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]


def test_balanced_shards_equal_counts_and_tokens():
    """trainer.balanced_shards: every rank gets B / world impressions, token totals within 1 % of each other, deterministic,
    and a partition of the global batch (same set of impressions as any other sharding)"""
    import numpy as np
    from nnr_b200.trainer import balanced_shards
    rng = np.random.default_rng(0)
    cost = rng.integers(200, 9000, size=512)
    shards = balanced_shards(cost, 8)
    assert [len(s) for s in shards] == [64] * 8
    assert sorted(i for s in shards for i in s) == list(range(512))
    totals = np.array([cost[s].sum() for s in shards])
    assert totals.max() / totals.mean() < 1.01
    assert shards == balanced_shards(cost, 8)
    random_totals = np.array([cost[r * 64:(r + 1) * 64].sum() for r in range(8)])
    assert totals.max() < random_totals.max()
    import pytest
    with pytest.raises(ValueError):
        balanced_shards(cost[:510], 8)
