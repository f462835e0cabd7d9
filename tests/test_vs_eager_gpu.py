"""The thing this path replaces ON THE GPU is stock PyTorch eager (the reference ships no kernels:
SURVEY.md section 2.1).  The oracle's op-for-op torch port issues the same ATen sequence as
patchnce_cut.py; this test runs both on the same device tensors, checks they agree, and requires the
hand-written path to be faster (it prints both timings: run with -s to see them)."""
import time

import pytest
import torch

pytestmark = pytest.mark.gpu

B5 = [(64, 256, 256), (256, 64, 64), (256, 64, 64), (128, 128, 128), (64, 256, 256)]


def _timed(fn, n):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


@pytest.mark.parametrize("b", [1, 8])
def test_faster_than_eager_reference_port(b):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gan_variant_research_b200 as pn
    from oracle import patchnce_oracle as orc
    g = torch.Generator(device="cuda").manual_seed(3)
    src = [torch.randn(b, *s, device="cuda", generator=g).relu() for s in B5]
    tgt = [torch.randn(b, *s, device="cuda", generator=g).relu().requires_grad_() for s in B5]
    crit = pn.PatchNCELoss(0.07, 256)

    def ours():
        for t in tgt:
            t.grad = None
        torch.manual_seed(7)
        loss = crit(src, tgt)
        loss.backward()
        return loss

    def eager():
        for t in tgt:
            t.grad = None
        torch.manual_seed(7)
        loss, _ = orc.patchnce_loss_torch(src, tgt, 0.07, 256)
        loss.backward()
        return loss

    lo = ours()
    go = [t.grad.clone() for t in tgt]
    le = eager()
    assert lo.item() == pytest.approx(le.item(), rel=1e-4)
    for a, t in zip(go, tgt):
        scale = t.grad.abs().max()
        assert (a - t.grad).abs().max() <= 2e-3 * scale     # eager uses TF32-free fp32 mm; ours bf16x3
    t_ours, t_eager = _timed(ours, 20), _timed(eager, 5)
    print(f"\nB={b}: ours {t_ours * 1e3:.3f} ms/step, eager port of the reference {t_eager * 1e3:.3f} ms/step, "
          f"x{t_eager / t_ours:.1f}")
    assert t_ours < t_eager
