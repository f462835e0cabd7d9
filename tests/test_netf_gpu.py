"""PatchSampleF(use_mlp=True).forward(feats, num_patches, patch_ids) -- the north-star module signature -- running in
libpnce on the tensor cores (pnce_netf_fwd / pnce_netf_bwd), against the oracle's torch restatement of the head
(PARITY UNPINNED by the reference, which has no netF head: SURVEY.md section 8 row a13).  Tolerance 1e-3 (north_star)."""
import numpy as np
import pytest
import torch

from test_parity_gpu import _head_problem, assert_grad_close

pytestmark = pytest.mark.gpu
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


def _make_head(pn, nc, feats, seed=5):
    torch.manual_seed(seed)
    netF = pn.PatchSampleF(use_mlp=True, nc=nc, init_gain=0.3)
    netF.create_mlp(feats)
    for prm in netF.parameters():
        if prm.dim() == 1:
            torch.nn.init.normal_(prm, 0.0, 0.1)
    return netF


def _cpu_heads(netF, n):
    heads = []
    for l in range(n):
        mlp = getattr(netF, f"mlp_{l}")
        heads.append(tuple(x.detach().cpu().clone().requires_grad_()
                           for x in (mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias)))
    return heads


def _oracle_rows(orc, feats, ids, heads):
    out = []
    for f, i, (w1, b1, w2, b2) in zip(feats, ids, heads):
        b, c = f.shape[:2]
        rows = f.reshape(b, c, -1).transpose(1, 2)[:, i, :].reshape(-1, c)
        out.append(orc.head_forward_torch(rows, w1, b1, w2, b2))
    return out


def _relu_boundary(feat, ids, w1, b1):
    """(row, unit) pairs whose hidden pre-activation is zero to within the rounding of the bf16x3 contraction
    (2^-16 of sum |terms|): ReLU's derivative is discontinuous there, so the tensor-core path and fp32 may mask that ONE
    unit differently -- it then enters or leaves d feat[row], dW1[unit] and db1[unit] (DESIGN.md section 2,
    scratch/stress2.py).  Returns (bool (B, P) rows, bool (nc,) units); they are left out of the comparison, everything
    else is held to the tolerance."""
    f = feat.detach().double()
    b, c = f.shape[:2]
    rows = f.reshape(b, c, -1).transpose(1, 2)[:, ids, :]
    w, bb = w1.detach().double(), b1.detach().double()
    pre = rows @ w.t() + bb
    mag = rows.abs() @ w.abs().t() + bb.abs()
    amb = pre.abs() <= 2.0 ** -16 * mag
    return amb.any(-1), amb.any(0).any(0)


def _drop_rows(grad, ids, amb_rows):
    """zero the sampled positions of ambiguous rows (numpy (B, C, H, W) copy)"""
    g = np.array(grad, dtype=np.float64)
    flat = g.reshape(g.shape[0], g.shape[1], -1)
    bs, ps = np.nonzero(amb_rows.numpy())
    flat[bs, :, np.asarray(ids)[ps]] = 0.0
    return g


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("nc,p", [(256, 64), (128, 200), (256, 300)])
def test_patch_sample_f_head_forward_and_backward(pn, orc, nc, p, layout):
    """Forward rows (B*P, nc) in ids order and, for a random cotangent, d feat (dense, exact zeros off the samples)
    and the four parameter gradients of every map.  The launch list of this call holds libpnce kernels only
    (k_prep, k_wprep, k_gather_tc, k_gemm_tc_p x2 | k_netf_dy_pack, k_gemm_tc_p x2, k_wgrad_tc, k_wreduce,
    k_dense_*): asserted below with the torch profiler."""
    shapes = [(64, 32, 32), (128, 16, 16), (24, 20, 12)]
    _, tgt, ids = _head_problem(321, 3, shapes, p)
    feats = [x.cuda() for x in tgt]
    if layout == "nhwc":
        feats = [x.contiguous(memory_format=CL) for x in feats]
    netF = _make_head(pn, nc, feats)
    t = [x.requires_grad_() for x in feats]
    idd = [i.cuda() for i in ids]
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        rows, rid = netF(t, p, idd)
        g = torch.Generator(device="cuda").manual_seed(1)
        cot = [torch.randn(r.shape, device="cuda", generator=g) for r in rows]
        torch.autograd.backward(rows, cot)
        torch.cuda.synchronize()
    names = {e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA}
    assert any("k_gemm_tc_p" in n for n in names) and any("k_wgrad_tc" in n for n in names)
    assert not any(("gemm" in n.lower() or "cutlass" in n.lower() or "cublas" in n.lower()) and "pnce" not in n for n in names), names
    assert pn.poll_nonfinite_warnings(block=True) == 0
    assert all(torch.equal(a, b) for a, b in zip(rid, idd))
    heads = _cpu_heads(netF, len(shapes))
    tc = [x.clone().requires_grad_() for x in tgt]
    want = _oracle_rows(orc, tc, ids, heads)
    torch.autograd.backward(want, [c.cpu() for c in cot])
    for l in range(len(shapes)):
        assert rows[l].shape == want[l].shape and rows[l].dtype == torch.float32
        assert_grad_close(rows[l].detach().cpu().numpy(), want[l].detach().numpy(), 1e-3, f"rows layer {l}")
        if layout == "nhwc" and shapes[l][0] > 1:
            assert t[l].grad.is_contiguous(memory_format=CL)
        amb_rows, amb_units = _relu_boundary(tgt[l], ids[l], heads[l][0], heads[l][1])
        assert amb_rows.float().mean() < 0.2
        assert_grad_close(_drop_rows(t[l].grad.cpu().numpy(), ids[l], amb_rows), _drop_rows(tc[l].grad.numpy(), ids[l], amb_rows),
                          1e-3, f"d feat layer {l}", ids=ids[l])
        mlp = getattr(netF, f"mlp_{l}")
        keep = (~amb_units).numpy()
        for got, w, name in zip((mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias), heads[l], ("W1", "b1", "W2", "b2")):
            a, c = got.grad.cpu().numpy(), w.grad.numpy()
            if name in ("W1", "b1"):
                a, c = a[keep], c[keep]
            assert_grad_close(a, c, 1e-3, f"d{name} layer {l}")


def test_patch_sample_f_head_no_grad_and_drawn_ids(pn, orc):
    """Source side of CUT: under no_grad, patch_ids=None (ids drawn like the reference, :60-63); rows follow the oracle."""
    shapes = [(64, 32, 32), (256, 16, 16)]
    _, tgt, _ = _head_problem(77, 2, shapes, 256)
    feats = [x.cuda() for x in tgt]
    netF = _make_head(pn, 256, feats)
    torch.manual_seed(3)
    with torch.no_grad():
        rows, ids = netF(feats, 256, None)
    torch.manual_seed(3)
    want_ids = [torch.randint(0, s[1] * s[2], (256,), device="cuda") for s in shapes]
    assert all(torch.equal(a, b) for a, b in zip(ids, want_ids))
    heads = _cpu_heads(netF, 2)
    with torch.no_grad():
        want = _oracle_rows(orc, tgt, [i.cpu() for i in ids], heads)
    for r, w in zip(rows, want):
        assert not r.requires_grad
        assert_grad_close(r.cpu().numpy(), w.numpy(), 1e-3, "rows")


@pytest.mark.parametrize("fused", [True, False])
def test_head_with_1024_patches(pn, orc, fused):
    """BASELINE config 4's patch count with the head: the fused call (key-blocked loss kernel on the head's output) and
    the module composition (PatchSampleF on tcgen05 + rows loss) against the oracle."""
    shapes = [(64, 48, 48), (128, 40, 40)]
    src, tgt, ids = _head_problem(55, 2, shapes, 1024)
    netF = _make_head(pn, 256, [x.cuda() for x in tgt])
    t = [x.cuda().requires_grad_() for x in tgt]
    idd = [i.cuda() for i in ids]
    loss, _ = pn.patchnce_with_head(netF, [x.cuda() for x in src], t, 0.07, 1024, idd, fused=fused)
    loss.backward()
    assert pn.poll_nonfinite_warnings(block=True) == 0
    heads = _cpu_heads(netF, len(shapes))
    tc = [x.clone().requires_grad_() for x in tgt]
    want = orc.patchnce_head_loss_torch(src, tc, ids, heads)
    want.backward()
    assert loss.item() == pytest.approx(want.item(), rel=1e-3)
    for l in range(len(shapes)):
        amb_rows, amb_units = _relu_boundary(tgt[l], ids[l], heads[l][0], heads[l][1])
        assert_grad_close(_drop_rows(t[l].grad.cpu().numpy(), ids[l], amb_rows), _drop_rows(tc[l].grad.numpy(), ids[l], amb_rows),
                          1e-3, f"d tgt layer {l}", ids=ids[l])
        mlp = getattr(netF, f"mlp_{l}")
        keep = (~amb_units).numpy()
        for got, w, name in zip((mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias), heads[l], ("W1", "b1", "W2", "b2")):
            a, c = got.grad.cpu().numpy(), w.grad.numpy()
            if name in ("W1", "b1"):
                a, c = a[keep], c[keep]
            assert_grad_close(a, c, 1e-3, f"d{name} layer {l}")


@pytest.mark.parametrize("p", [64, 300])
def test_fused_head_on_channels_last_maps_equals_nchw(pn, p):
    """patchnce_with_head (fused) on torch.channels_last maps: same loss, dense gradient (channels-last, exact zeros off
    the samples) and head gradients as on the NCHW storage of the same values (same kernels after the gather, so the
    agreement is to rounding of the norms, not 1e-3)."""
    shapes = [(64, 32, 32), (128, 16, 16), (24, 20, 12)]
    src, tgt, ids = _head_problem(11, 3, shapes, p)
    netF = _make_head(pn, 256, [x.cuda() for x in tgt])
    idd = [i.cuda() for i in ids]
    out = []
    for cl in (False, True):
        netF.zero_grad()
        mk = (lambda x: x.cuda().contiguous(memory_format=CL)) if cl else (lambda x: x.cuda())
        t = [mk(x).requires_grad_() for x in tgt]
        loss, _ = pn.patchnce_with_head(netF, [mk(x) for x in src], t, 0.07, p, idd, fused=True)
        loss.backward()
        out.append((loss.detach(), [x.grad for x in t], [q.grad.clone() for q in netF.parameters()]))
    assert pn.poll_nonfinite_warnings(block=True) == 0
    (l0, g0, w0), (l1, g1, w1) = out
    assert l1.item() == pytest.approx(l0.item(), rel=1e-6)
    for a, c, i in zip(g0, g1, ids):
        if a.shape[1] > 1:
            assert c.is_contiguous(memory_format=CL)
        assert_grad_close(c.cpu().numpy(), a.cpu().numpy(), 1e-5, "d tgt", ids=i)
    for a, c in zip(w0, w1):
        assert_grad_close(c.cpu().numpy(), a.cpu().numpy(), 1e-5, "head gradient")
