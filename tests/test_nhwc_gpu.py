"""Channels-last (NHWC) fast path: the same loss law on torch.channels_last maps (an extension of the reference's
interface -- its `.view(B, C, -1)`, patchnce_cut.py:56, rejects such tensors).  The oracle and the reference-frozen
fixtures are layout-agnostic (they see the logical (B, C, H, W) values), so parity = same fixtures, same tolerances,
inputs merely stored (B, H, W, C); the gradient must come back channels-last with exact zeros off the samples."""
import glob
import os

import numpy as np
import pytest
import torch

from test_parity_gpu import B5, MATH_TOL, assert_grad_close, load_small

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = sorted(glob.glob(os.path.join(HERE, "golden", "small_*.npz")))
CL = torch.channels_last


@pytest.fixture(scope="module")
def pn():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gan_variant_research_b200 as m
    from gan_variant_research_b200 import _lib
    _lib.load()
    return m


@pytest.fixture(scope="module")
def orc():
    from oracle import patchnce_oracle
    return patchnce_oracle


def cl(x):
    """numpy (B, C, H, W) -> CUDA tensor of the same logical values in channels-last storage"""
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda().contiguous(memory_format=CL)
    return t


def is_cl(t):
    return t.is_contiguous(memory_format=CL)


@pytest.mark.parametrize("math", ["tc_bf16x3", "tc_bf16"])
@pytest.mark.parametrize("path", SMALL, ids=[os.path.basename(p)[6:-4] for p in SMALL])
def test_channels_last_matches_reference_goldens(pn, path, math):
    d, src, tgt, ids, grads = load_small(path)
    t = [cl(x).requires_grad_() for x in tgt]
    loss = pn.fused_patchnce([cl(x) for x in src], t, [torch.from_numpy(i).cuda() for i in ids], float(d["tau"]), math=math)
    (loss * float(d["upstream"])).backward()
    ltol, gtol = MATH_TOL[math]
    if math == "tc_bf16" and float(d["tau"]) < 0.05:
        ltol, gtol = 2e-2, 2e-1
    assert loss.item() == pytest.approx(float(d["loss"]), rel=ltol, abs=1e-6)
    for i, (tt, g) in enumerate(zip(t, grads)):
        assert tt.grad.shape == tt.shape
        if tt.shape[1] > 1 and tt.shape[2] * tt.shape[3] > 1:
            assert is_cl(tt.grad), "the gradient of a channels-last map is channels-last"
        assert_grad_close(tt.grad.cpu().numpy(), g, gtol, f"layer {i}", ids=ids[i])
    assert pn.poll_nonfinite_warnings(block=True) >= 0


@pytest.mark.parametrize("b,c,h,w,p", [(5, 200, 20, 20, 200), (1, 130, 16, 16, 129), (7, 64, 12, 12, 256),
                                       (3, 255, 17, 15, 255), (2, 32, 128, 128, 160), (2, 3, 9, 9, 50),
                                       (2, 96, 24, 24, 300), (1, 256, 40, 40, 1024)])
def test_channels_last_odd_shapes_against_the_oracle(pn, orc, b, c, h, w, p):
    """Channel counts that are not multiples of 8 / 32 (scalar loads, scalar row stores, tiles of a few odd-sized
    positions), duplicate ids, P > 256 (the key-blocked kernel's row-major epilogue)."""
    g = torch.Generator().manual_seed(b * 1000 + c)
    src = [torch.randn(b, c, h, w, generator=g)]
    tgt = [torch.randn(b, c, h, w, generator=g)]
    ids = [torch.randint(0, h * w, (min(p, h * w),), generator=g)]
    t = [x.cuda().contiguous(memory_format=CL).requires_grad_() for x in tgt]
    loss = pn.fused_patchnce([x.cuda().contiguous(memory_format=CL) for x in src], t, [i.cuda() for i in ids], 0.07)
    (loss * 0.5).backward()
    want, _, gw = orc.patchnce_loss_and_grads_np([x.numpy() for x in src], [x.numpy() for x in tgt],
                                                 [i.numpy() for i in ids], 0.07, upstream=0.5)
    assert loss.item() == pytest.approx(want, rel=2e-5)
    assert is_cl(t[0].grad)
    assert_grad_close(t[0].grad.cpu().numpy(), gw[0], 2e-4, "odd shape", ids=ids[0].numpy())
    assert pn.poll_nonfinite_warnings(block=True) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_channels_last_equals_nchw_at_full_size(pn, dtype):
    """B5 maps at B=3 through PatchNCELoss.forward (ids drawn by the library): channels-last and NCHW storage of the
    same values give the same ids, the same loss (different summation order of the norms: 1e-6) and the same dense
    gradient; the source maps may come in the other layout (they are re-laid out to match the target's)."""
    b = 3
    g = torch.Generator(device="cuda").manual_seed(5)
    src = [torch.randn(b, *s, device="cuda", generator=g).relu().to(dtype) for s in B5]
    tgt = [torch.randn(b, *s, device="cuda", generator=g).relu().to(dtype) for s in B5]
    crit = pn.PatchNCELoss(0.07, 256, [0, 4, 8, 12, 13])
    torch.manual_seed(11)
    t0 = [x.clone().requires_grad_() for x in tgt]
    l0 = crit(src, t0)
    ids0 = [i.clone() for i in crit.last_patch_ids]
    l0.backward()
    torch.manual_seed(11)
    t1 = [x.contiguous(memory_format=CL).requires_grad_() for x in tgt]
    s1 = [x.contiguous(memory_format=CL) if k % 2 == 0 else x for k, x in enumerate(src)]
    l1 = crit(s1, t1)
    for a, c in zip(ids0, crit.last_patch_ids):
        assert torch.equal(a, c)
    l1.backward()
    assert l1.item() == pytest.approx(l0.item(), rel=2e-6 if dtype == torch.float32 else 1e-5)
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    for a, c, i in zip(t0, t1, ids0):
        assert is_cl(c.grad) and c.grad.dtype == dtype
        scale = a.grad.float().abs().max().item()
        assert (a.grad.float() - c.grad.float()).abs().max().item() <= tol * scale
        mask = torch.zeros(a.shape[2] * a.shape[3], dtype=torch.bool, device="cuda")
        mask[i] = True
        assert not (c.grad.reshape(b, a.shape[1], -1)[:, :, ~mask] != 0).any()
    assert pn.poll_nonfinite_warnings(block=True) == 0


def test_channels_last_loss_and_grads_and_upstream(pn):
    """The autograd-free entry on channels-last maps, with an upstream gradient: equals forward + backward."""
    b = 2
    g = torch.Generator(device="cuda").manual_seed(6)
    src = [torch.randn(b, *s, device="cuda", generator=g).contiguous(memory_format=CL) for s in B5[1:4]]
    tgt = [torch.randn(b, *s, device="cuda", generator=g).contiguous(memory_format=CL) for s in B5[1:4]]
    crit = pn.PatchNCELoss(0.07, 256)
    torch.manual_seed(3)
    t = [x.clone(memory_format=torch.preserve_format).requires_grad_() for x in tgt]
    loss = crit(src, t)
    loss.backward(torch.tensor(3.0, device="cuda"))
    torch.manual_seed(3)
    loss2, grads = crit.loss_and_grads(src, tgt, torch.tensor(3.0, device="cuda"))
    assert torch.equal(loss, loss2)
    for a, c in zip(t, grads):
        assert is_cl(c)
        assert torch.equal(a.grad, c)


def test_channels_last_outside_the_tensor_core_envelope_is_relaid_out(pn, orc):
    """C > 256 or math='simt_f32': the maps are re-laid out to NCHW (as any non-contiguous input is) and the fp32
    CUDA-core kernels run; values still follow the oracle."""
    g = torch.Generator().manual_seed(8)
    src = [torch.randn(2, 300, 10, 10, generator=g)]
    tgt = [torch.randn(2, 300, 10, 10, generator=g)]
    ids = [torch.randint(0, 100, (64,), generator=g)]
    want, _, gw = orc.patchnce_loss_and_grads_np([x.numpy() for x in src], [x.numpy() for x in tgt],
                                                 [i.numpy() for i in ids], 0.07)
    for math in ("tc_bf16x3", "simt_f32"):
        t = [x.cuda().contiguous(memory_format=CL).requires_grad_() for x in tgt]
        loss = pn.fused_patchnce([x.cuda().contiguous(memory_format=CL) for x in src], t, [i.cuda() for i in ids], 0.07,
                                 math=math)
        loss.backward()
        assert loss.item() == pytest.approx(want, rel=2e-5)
        assert_grad_close(t[0].grad.cpu().numpy(), gw[0], 2e-4, math, ids=ids[0].numpy())


def test_channels_last_pinned_host_maps_zero_copy(pn):
    """Channels-last maps left in pinned HOST memory (pinned_as_device aliases the dense (B, H, W, C) storage): the gather
    pulls each sampled patch as one row over PCIe; results are bit-identical to device-resident channels-last copies."""
    g = torch.Generator().manual_seed(8)
    shapes = [(2, 32, 24, 24), (2, 64, 16, 16)]
    def pinned_cl(s):
        b, c, h, w = s
        base = torch.randn(b, h, w, c, generator=g).pin_memory()
        return base.permute(0, 3, 1, 2)
    h_src = [pinned_cl(s) for s in shapes]
    h_tgt = [pinned_cl(s) for s in shapes]
    a_src = [pn.pinned_as_device(h) for h in h_src]
    a_tgt = [pn.pinned_as_device(h).requires_grad_() for h in h_tgt]
    assert a_src[0].is_cuda and a_src[0].data_ptr() == h_src[0].data_ptr() and is_cl(a_src[0]) and a_src[0].shape == h_src[0].shape
    d_src = [h.cuda() for h in h_src]
    d_tgt = [h.cuda().requires_grad_() for h in h_tgt]
    assert is_cl(d_src[0]) and not d_src[0].is_contiguous()
    crit = pn.PatchNCELoss(0.07, 128)
    torch.manual_seed(5)
    la = crit(a_src, a_tgt)
    la.backward()
    torch.manual_seed(5)
    ld = crit(d_src, d_tgt)
    ld.backward()
    assert la.item() == ld.item()
    for x, y in zip(a_tgt, d_tgt):
        assert x.grad.is_cuda and torch.equal(x.grad, y.grad)
