"""D-side row.  CPU: the numpy oracle, fed with draws replayed in the order the B200 host
makes them, against the fixture frozen from the reference's DiffAugment and hinge losses.  GPU: pnce_diffaug /
pnce_hinge_* against the fixture and the oracle, RNG stream alignment, half precision, error behaviour."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "dside_reference.npz")
POLICIES = [["color", "translation", "cutout"], ["color", "translation", "cutout_light"], ["translation"], ["color"],
            ["cutout"], ["translation", "cutout_light"]]
SHAPE = (3, 3, 20, 24)


def inputs():
    g = torch.Generator().manual_seed(51)
    return torch.randn(SHAPE, generator=g), torch.randn(SHAPE, generator=g)


def params(d, k):
    color = [d[f"aug{k}:color{i}"] for i in range(3)] if f"aug{k}:color0" in d else None
    shift = [d[f"aug{k}:shift{i}"] for i in range(2)] if f"aug{k}:shift0" in d else None
    cut = [d[f"aug{k}:cut{i}"] for i in range(2)] if f"aug{k}:cut0" in d else None
    cut_hw = tuple(int(v) for v in d[f"aug{k}:cut_hw"]) if cut is not None else (0, 0)
    return color, shift, cut, cut_hw


def close(got, want, tol, what=""):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, what
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
    assert err <= tol, f"{what}: {err:.2e} > {tol}"


@pytest.mark.parametrize("k", range(len(POLICIES)), ids=["+".join(p) for p in POLICIES])
def test_oracle_matches_the_reference_fixture(k):
    from oracle import dside_oracle as orc
    d = np.load(GOLD)
    x, up = inputs()
    color, shift, cut, cut_hw = params(d, k)
    assert (color is not None) == ("color" in POLICIES[k]) and (shift is not None) == ("translation" in POLICIES[k])
    close(orc.diffaug_np(x.numpy(), color, shift, cut, cut_hw), d[f"aug{k}:y"], 1e-6, "forward")
    close(orc.diffaug_vjp_np(up.numpy(), color, shift, cut, cut_hw), d[f"aug{k}:dx"], 1e-6, "backward")
    # exact zeros where the reference has them (padding, cut-out box)
    y = orc.diffaug_np(x.numpy(), color, shift, cut, cut_hw)
    assert np.array_equal(y == 0, d[f"aug{k}:y"] == 0)


def test_hinge_oracle_matches_the_reference_fixture():
    from oracle import dside_oracle as orc
    d = np.load(GOLD)
    real = [d[f"hinge:real{i}"].astype(np.float64) for i in range(2)]
    fake = [d[f"hinge:fake{i}"].astype(np.float64) for i in range(2)]
    loss, dr, df = orc.d_hinge_np(real, fake)
    assert loss == pytest.approx(float(d["hinge:d_loss"]), rel=1e-6)
    for i in range(2):
        close(3.0 * dr[i], d[f"hinge:d_dreal{i}"], 1e-6); close(3.0 * df[i], d[f"hinge:d_dfake{i}"], 1e-6)
    loss, df = orc.g_hinge_np(fake)
    assert loss == pytest.approx(float(d["hinge:g_loss"]), rel=1e-6)
    for i in range(2):
        close(3.0 * df[i], d[f"hinge:g_dfake{i}"], 1e-6)


def test_hinge_oracle_propagates_nonfinite_scores_like_the_reference():
    """A diverged discriminator (NaN / Inf scores): d_loss must be NaN so that train_step's check raises
    (train_cutpp.py:326-329); relu's backward passes the gradient at a NaN input (fixture from the reference)."""
    from oracle import dside_oracle as orc
    d = np.load(GOLD)
    real = [d[f"hinge_nan:real{i}"].astype(np.float64) for i in range(2)]
    fake = [d[f"hinge_nan:fake{i}"].astype(np.float64) for i in range(2)]
    with np.errstate(invalid="ignore"):
        loss, dr, df = orc.d_hinge_np(real, fake)
    assert np.isnan(loss) and np.isnan(float(d["hinge_nan:d_loss"]))
    for i in range(2):
        close(3.0 * dr[i], d[f"hinge_nan:d_dreal{i}"], 1e-6); close(3.0 * df[i], d[f"hinge_nan:d_dfake{i}"], 1e-6)


@pytest.mark.gpu
def test_cuda_hinge_propagates_nonfinite_scores_like_the_reference():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gan_variant_research_b200 as pn
    d = np.load(GOLD)
    real = [_dev(d[f"hinge_nan:real{i}"], torch.float32).requires_grad_() for i in range(2)]
    fake = [_dev(d[f"hinge_nan:fake{i}"], torch.float32).requires_grad_() for i in range(2)]
    ld = pn.discriminator_hinge_loss(real, fake)
    (ld * 3.0).backward()
    assert np.isnan(ld.item())
    for i in range(2):
        close(real[i].grad.cpu().numpy(), d[f"hinge_nan:d_dreal{i}"], 2e-6, "d real")
        close(fake[i].grad.cpu().numpy(), d[f"hinge_nan:d_dfake{i}"], 2e-6, "d fake")


def test_policy_validation_and_cpu_tensors_fail_loudly():
    import gan_variant_research_b200 as pn
    with pytest.raises(NotImplementedError):
        pn.DiffAugment(["translation", "color"])
    with pytest.raises(NotImplementedError):
        pn.DiffAugment(["cutout", "cutout_light"])
    aug = pn.DiffAugment()                                            # default policy of the reference (:72-73)
    assert aug.policy == ["color", "translation", "cutout_light"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        aug(torch.zeros(1, 3, 8, 8))
    assert pn.DiffAugment(["nonsense"])(torch.zeros(1, 3, 8, 8)).shape == (1, 3, 8, 8)   # unknown names are skipped (:77)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pn.generator_hinge_loss(torch.zeros(2, 1, 3, 3))


# ------------------------------------------------------------------------------------------------------------------
def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float16, 3e-3), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("k", range(len(POLICIES)), ids=["+".join(p) for p in POLICIES])
def test_cuda_diffaug_matches_fixture(k, dtype, tol):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gan_variant_research_b200 import dside
    d = np.load(GOLD)
    x, up = inputs()
    color, shift, cut, cut_hw = params(d, k)
    b = SHAPE[0]
    col = tuple(_dev(c, dtype).reshape(b, 1, 1, 1) for c in color) if color else None
    sh = tuple(_dev(s).reshape(b, 1, 1) for s in shift) if shift else None
    ct = tuple(_dev(c).reshape(b, 1, 1) for c in cut) if cut else None
    xd = x.cuda().to(dtype).requires_grad_()
    y = dside._DiffAugFn.apply(xd, col, sh, ct, cut_hw)
    assert y.dtype == dtype and y.shape == xd.shape
    y.backward(up.cuda().to(dtype))
    close(y.detach().float().cpu().numpy(), d[f"aug{k}:y"], tol, "forward")
    close(xd.grad.float().cpu().numpy(), d[f"aug{k}:dx"], tol, "backward")
    if dtype == torch.float32:
        assert np.array_equal(y.detach().cpu().numpy() == 0, d[f"aug{k}:y"] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("policy", [["color", "translation", "cutout"], ["translation", "cutout_light"], ["color"]])
def test_cuda_diffaug_draws_like_the_reference(policy):
    """Same seed -> the module draws what the reference's functions would draw on this device, in the same order, and
    leaves the generator where they would leave it."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gan_variant_research_b200 as pn
    from oracle import dside_oracle as orc
    b, c, h, w = 4, 3, 64, 48
    x = torch.randn(b, c, h, w, device="cuda")
    torch.manual_seed(123)
    y = pn.DiffAugment(policy)(x)
    after = torch.rand(3, device="cuda")
    torch.manual_seed(123)
    color = shift = cut = None
    cut_hw = (0, 0)
    if "color" in policy:                                             # diffaugment.py:8, :15, :22
        color = [torch.rand(b, 1, 1, 1, dtype=x.dtype, device=x.device).cpu().numpy().reshape(b) for _ in range(3)]
    if "translation" in policy:                                       # :27-29
        sx, sy = int(h * 0.125 + 0.5), int(w * 0.125 + 0.5)
        shift = [torch.randint(-sx, sx + 1, size=[b, 1, 1], device=x.device).cpu().numpy().reshape(b),
                 torch.randint(-sy, sy + 1, size=[b, 1, 1], device=x.device).cpu().numpy().reshape(b)]
    ratio = 0.5 if "cutout" in policy else (0.2 if "cutout_light" in policy else None)
    if ratio is not None:                                             # :45-47
        cut_hw = (int(h * ratio + 0.5), int(w * ratio + 0.5))
        cut = [torch.randint(0, h + (1 - cut_hw[0] % 2), size=[b, 1, 1], device=x.device).cpu().numpy().reshape(b),
               torch.randint(0, w + (1 - cut_hw[1] % 2), size=[b, 1, 1], device=x.device).cpu().numpy().reshape(b)]
    assert torch.equal(after, torch.rand(3, device="cuda"))           # generator state aligned
    close(y.cpu().numpy(), orc.diffaug_np(x.cpu().numpy(), color, shift, cut, cut_hw), 2e-6, "vs oracle")


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float16, 2e-3)])
def test_cuda_hinge_matches_fixture(dtype, tol):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gan_variant_research_b200 as pn
    d = np.load(GOLD)
    real = [_dev(d[f"hinge:real{i}"], dtype).requires_grad_() for i in range(2)]
    fake = [_dev(d[f"hinge:fake{i}"], dtype).requires_grad_() for i in range(2)]
    ld = pn.discriminator_hinge_loss(real, fake)
    (ld * 3.0).backward()
    assert ld.item() == pytest.approx(float(d["hinge:d_loss"]), rel=max(tol, 1e-6))
    for i in range(2):
        close(real[i].grad.float().cpu().numpy(), d[f"hinge:d_dreal{i}"], tol, "d real")
        close(fake[i].grad.float().cpu().numpy(), d[f"hinge:d_dfake{i}"], tol, "d fake")
        fake[i].grad = None
    lg = pn.generator_hinge_loss(fake)
    (lg * 3.0).backward()
    assert lg.item() == pytest.approx(float(d["hinge:g_loss"]), rel=max(tol, 1e-6))
    for i in range(2):
        close(fake[i].grad.float().cpu().numpy(), d[f"hinge:g_dfake{i}"], tol, "g d fake")
    # single tensors instead of lists (adv_hinge.py:19-21, :47-48); only the fake side needs a gradient in the D step
    r, f = real[0].detach(), fake[0].detach().requires_grad_()
    l1 = pn.discriminator_hinge_loss(r, f)
    l1.backward()
    want = 0.5 * (torch.relu(1 - r.float()).mean() + torch.relu(1 + f.float()).mean())
    assert l1.item() == pytest.approx(want.item(), rel=max(tol, 1e-6)) and f.grad is not None
