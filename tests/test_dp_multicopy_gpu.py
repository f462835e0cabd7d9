"""dp._MultiCopy: the GradReducer's bucket <-> gradient copies as one libpnce launch each way (single GPU; the
two-GPU reducer test in test_dp_nccl_gpu.py drives it through NCCL)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_bucket_pack_and_unpack_are_exact_and_follow_moved_gradients():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gan_variant_research_b200 import dp
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    shapes = [(3,), (16, 3, 3, 3), (1,), (70001,), (256, 256, 3, 3), (5, 7)]       # odd sizes: padded slots, scalar tails
    params = [torch.zeros(*s, device=dev).requires_grad_() for s in shapes]
    slots, off = [], 0
    for p in params:
        slots.append((p, off, p.numel()))
        off += (p.numel() + 3) // 4 * 4
    flat = torch.full((off,), -7.0, device=dev)
    mc = dp._MultiCopy(flat, slots)
    assert not mc.pack()                                                           # no gradients yet: caller falls back
    for rnd in range(2):                                                           # second round: every gradient tensor is new
        for p in params:
            p.grad = torch.randn(p.shape, device=dev, generator=g)
        assert mc.pack()
        for p, o, n in slots:
            assert torch.equal(flat[o:o + n], p.grad.reshape(-1))
        flat.mul_(0.5)
        want = [p.grad * 0.5 for p in params]
        versions = [p.grad._version for p in params]
        assert mc.unpack()
        for p, w, v in zip(params, want, versions):
            assert torch.equal(p.grad, w)
            assert p.grad._version > v                                             # autograd sees the in-place write
    params[1].grad = torch.randn(3, 3, 3, 16, device=dev, generator=g).permute(3, 2, 1, 0)   # same shape, not contiguous:
    assert not params[1].grad.is_contiguous()                                      # the caller's per-slot copies take over
    assert not mc.pack() and not mc.unpack()
