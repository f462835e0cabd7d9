"""Data-parallel host logic on CPU: world_size 2, gloo.  The compute inside
each rank is the oracle's torch restatement of the head path (the CUDA path cannot run here); what is
under test is the sharding / id agreement / flat all-reduce plumbing of gan_variant_research_b200.dp."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_problem():
    g = torch.Generator().manual_seed(3)
    shapes = [(12, 8, 8), (20, 6, 6)]
    src = [torch.randn(4, *s, generator=g) for s in shapes]
    tgt = [torch.randn(4, *s, generator=g) for s in shapes]
    ids = [torch.randint(0, s[1] * s[2], (16,), generator=g) for s in shapes]
    heads = []
    for s in shapes:
        heads.append(tuple(torch.randn(*sz, generator=g) * 0.2 for sz in ((16, s[0]), (16,), (16, 16), (16,))))
    return src, tgt, ids, heads


class _Heads(torch.nn.Module):
    def __init__(self, heads):
        super().__init__()
        self.params = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for h in heads for t in h])

    def groups(self):
        p = list(self.params)
        return [tuple(p[i:i + 4]) for i in range(0, len(p), 4)]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from gan_variant_research_b200 import dp
        from oracle import patchnce_oracle as orc
        src, tgt, ids, heads = _make_problem()
        # rank 1 starts from different ids: the broadcast must make them agree with rank 0
        my_ids = [i.clone() if rank == 0 else torch.zeros_like(i) for i in ids]
        dp.broadcast_patch_ids(my_ids)
        assert all(torch.equal(a, b) for a, b in zip(my_ids, ids))
        net = _Heads(heads)
        s_loc = dp.shard_batch(src)
        t_loc = [t.clone().requires_grad_() for t in dp.shard_batch(tgt)]
        assert s_loc[0].shape[0] == 4 // world
        loss = orc.patchnce_head_loss_torch(s_loc, t_loc, my_ids, net.groups())
        loss.backward()
        n = dp.allreduce_head_grads(net)
        assert n == sum(p.numel() for p in net.parameters())
        # the flat in-place reduction the fused head backward uses (mean over the group)
        assert dp.resolve_group(True) is dist.group.WORLD
        flat = torch.full((5,), float(rank + 1))
        dp.allreduce_flat_(flat, dp.resolve_group(True), average=True)
        assert torch.equal(flat, torch.full((5,), 1.5))
        lsum = loss.detach().clone()
        dist.all_reduce(lsum)
        torch.save({"loss": lsum / world, "head": [p.grad.clone() for p in net.parameters()],
                    "dtgt": [t.grad.clone() / world for t in t_loc]}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_data_parallel_equals_full_batch(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import patchnce_oracle as orc
    src, tgt, ids, heads = _make_problem()
    net = _Heads(heads)
    t_full = [t.clone().requires_grad_() for t in tgt]
    full = orc.patchnce_head_loss_torch(src, t_full, ids, net.groups())
    full.backward()
    r = [torch.load(os.path.join(tmp_path, f"r{k}.pt")) for k in range(world)]
    assert r[0]["loss"].item() == pytest.approx(full.item(), rel=1e-6)
    for k in range(world):          # every rank holds the same, full-batch head gradients
        for got, p in zip(r[k]["head"], net.parameters()):
            torch.testing.assert_close(got, p.grad, rtol=1e-5, atol=1e-8)
    for l in range(len(tgt)):       # dense feature gradients stay local to the owning rank
        got = torch.cat([r[k]["dtgt"][l] for k in range(world)])
        torch.testing.assert_close(got, t_full[l].grad, rtol=1e-5, atol=1e-9)


def test_single_process_helpers_are_no_ops():
    from gan_variant_research_b200 import dp
    x = torch.arange(8.0).reshape(4, 2)
    assert torch.equal(dp.shard_batch([x])[0], x)
    assert torch.equal(dp.shard_batch([x], rank=1, world=2)[0], x[2:])
    with pytest.raises(ValueError):
        dp.shard_batch([x], rank=0, world=3)
    lin = torch.nn.Linear(2, 2)
    lin(x).sum().backward()
    assert dp.allreduce_head_grads(lin) == 0
    ids = [torch.arange(3)]
    assert dp.broadcast_patch_ids(ids) is ids
    assert dp.resolve_group(True) is None            # no process group: nothing to reduce over
    flat = torch.ones(3)
    assert dp.allreduce_flat_(flat) is flat and torch.equal(flat, torch.ones(3))


def _reducer_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from gan_variant_research_b200 import dp
        torch.manual_seed(0)                                   # same weights on both ranks
        g_net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16),
                                    torch.nn.ReLU(), torch.nn.Linear(16, 3))
        d_net = torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.ReLU(), torch.nn.Linear(8, 1))
        params = list(g_net.parameters()) + list(d_net.parameters())
        red = dp.GradReducer(params, bucket_bytes=600)         # tiny buckets: several of them, some partial
        assert len(red.buckets) >= 3
        data = torch.Generator().manual_seed(5)
        x_all = torch.randn(8, 6, generator=data)
        x = dp.shard_batch([x_all])[0]
        out = {}
        # "D step": only d_net gets gradients (its input is detached) -> a subset of the buckets fills
        d_loss = d_net(g_net(x).detach()).mean()
        d_loss.backward()
        out["d_step"] = [None if p.grad is None else p.grad.clone() for p in params]
        for p in params:
            p.grad = None
        # "G step": gradients flow through both nets; a GradScaler-like factor rides along
        g_loss = (d_net(g_net(x)).mean() + g_net(x).pow(2).mean()) * 1024.0
        g_loss.backward()
        out["g_step"] = [p.grad.clone() for p in params]
        # second backward into existing grads (accumulation) stays consistent
        (g_net(x).sum() * 0.5).backward()
        out["accum"] = [p.grad.clone() for p in g_net.parameters()]
        red.remove()
        torch.save(out, os.path.join(out_dir, f"red{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_grad_reducer_matches_full_batch(tmp_path):
    """GradReducer on two gloo ranks: after every backward() the gradients equal the full-batch ones
    (mean losses over equal shards), including a pass that touches only part of the parameters."""
    world = 2
    mp.spawn(_reducer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    torch.manual_seed(0)
    g_net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16),
                                torch.nn.ReLU(), torch.nn.Linear(16, 3))
    d_net = torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.ReLU(), torch.nn.Linear(8, 1))
    params = list(g_net.parameters()) + list(d_net.parameters())
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(5))
    r = [torch.load(os.path.join(tmp_path, f"red{k}.pt")) for k in range(world)]
    d_net(g_net(x).detach()).mean().backward()
    for k in range(world):
        for got, p in zip(r[k]["d_step"], params):
            if p.grad is None:
                assert got is None
            else:
                torch.testing.assert_close(got, p.grad, rtol=1e-5, atol=1e-7)
    for p in params:
        p.grad = None
    ((d_net(g_net(x)).mean() + g_net(x).pow(2).mean()) * 1024.0).backward()
    for k in range(world):
        for got, p in zip(r[k]["g_step"], params):
            torch.testing.assert_close(got, p.grad, rtol=1e-5, atol=1e-5)
    # the accumulation pass: sum over the local shard, averaged over ranks = half the full-batch sum,
    # added to the (already averaged) gradients of the previous pass
    want = [p.grad.clone() for p in g_net.parameters()]
    for p in params:
        p.grad = None
    (g_net(x).sum() * 0.5).backward()
    for k in range(world):
        for got, w, p in zip(r[k]["accum"], want, g_net.parameters()):
            torch.testing.assert_close(got, (w + p.grad / world) / 1.0, rtol=1e-4, atol=1e-4)
