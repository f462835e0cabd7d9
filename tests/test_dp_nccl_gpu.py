"""Two ranks, two GPUs, NCCL: the fused head path with ``dp_group`` (head gradients averaged by one flat
all-reduce started on a side stream under the dense kernel) equals the full batch on one GPU.
Skipped on boxes with a single GPU (run it with ``gpurun --gpus 2``)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    g = torch.Generator().manual_seed(13)
    shapes = [(32, 16, 16), (64, 12, 12)]
    src = [torch.randn(4, *s, generator=g) for s in shapes]
    tgt = [torch.randn(4, *s, generator=g) for s in shapes]
    ids = [torch.randint(0, s[1] * s[2], (64,), generator=g) for s in shapes]
    return src, tgt, ids


def _run(pn, src, tgt, ids, dev, dp_group):
    torch.manual_seed(21)                               # identical head weights everywhere
    netF = pn.PatchSampleF(use_mlp=True, nc=128).to(dev)
    t = [x.to(dev).requires_grad_() for x in tgt]
    netF.create_mlp(t)
    loss, _ = pn.patchnce_with_head(netF, [x.to(dev) for x in src], t, 0.07, 64, patch_ids=[i.to(dev) for i in ids],
                                    dp_group=dp_group)
    loss.backward()
    torch.cuda.synchronize(dev)
    return loss.detach().cpu(), [p.grad.cpu() for p in netF.parameters()], [x.grad.cpu() for x in t]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import gan_variant_research_b200 as pn
        src, tgt, ids = _problem()
        per = src[0].shape[0] // world
        sl = slice(rank * per, (rank + 1) * per)
        loss, head, dtgt = _run(pn, [x[sl] for x in src], [x[sl] for x in tgt], ids, dev, True)
        torch.save({"loss": loss, "head": head, "dtgt": dtgt}, os.path.join(out_dir, f"r{rank}.pt"))
        assert pn.poll_nonfinite_warnings(block=True) == 0
    finally:
        dist.destroy_process_group()


def test_two_gpu_head_gradients_are_averaged_under_the_dense_kernel(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import gan_variant_research_b200 as pn
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    src, tgt, ids = _problem()
    full_loss, full_head, full_dtgt = _run(pn, src, tgt, ids, torch.device("cuda", 0), None)
    r = [torch.load(os.path.join(tmp_path, f"r{k}.pt")) for k in range(world)]
    assert (r[0]["loss"].item() + r[1]["loss"].item()) / 2 == pytest.approx(full_loss.item(), rel=1e-5)
    for k in range(world):          # every rank holds the full-batch head gradients
        for got, want in zip(r[k]["head"], full_head):
            scale = max(want.abs().max().item(), 1e-30)
            assert (got - want).abs().max().item() / scale < 5e-4
    for l in range(len(tgt)):       # dense feature gradients stay local; local mean -> 1 / world of the global one
        got = torch.cat([r[k]["dtgt"][l] for k in range(world)]) / world
        scale = max(full_dtgt[l].abs().max().item(), 1e-30)
        assert (got - full_dtgt[l]).abs().max().item() / scale < 5e-4


def _reducer_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from gan_variant_research_b200 import dp
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, padding=1), torch.nn.ReLU(),
                                  torch.nn.Conv2d(16, 16, 3, padding=1), torch.nn.ReLU(),
                                  torch.nn.Conv2d(16, 3, 3, padding=1)).to(dev)
        red = dp.GradReducer(net.parameters(), bucket_bytes=4096)
        x = torch.randn(4, 3, 16, 16, generator=torch.Generator().manual_seed(3))
        xs = dp.shard_batch([x])[0].to(dev)
        with torch.amp.autocast("cuda"):
            loss = net(xs).float().pow(2).mean()
        (loss * 256.0).backward()
        torch.cuda.synchronize(dev)
        torch.save([p.grad.cpu() for p in net.parameters()], os.path.join(out_dir, f"g{rank}.pt"))
        red.remove()
    finally:
        dist.destroy_process_group()


def test_two_gpu_grad_reducer_matches_full_batch(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mp.spawn(_reducer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, padding=1), torch.nn.ReLU(),
                              torch.nn.Conv2d(16, 16, 3, padding=1), torch.nn.ReLU(),
                              torch.nn.Conv2d(16, 3, 3, padding=1)).to(dev)
    x = torch.randn(4, 3, 16, 16, generator=torch.Generator().manual_seed(3)).to(dev)
    with torch.amp.autocast("cuda"):
        loss = net(x).float().pow(2).mean()
    (loss * 256.0).backward()
    for k in range(world):
        got = torch.load(os.path.join(tmp_path, f"g{k}.pt"))
        for g, p in zip(got, net.parameters()):
            scale = max(p.grad.abs().max().item(), 1e-30)
            assert (g - p.grad.cpu()).abs().max().item() / scale < 5e-3      # fp16 autocast convolutions
