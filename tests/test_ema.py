"""EMA row: the oracle against the fixture frozen from the reference's
utils/io_ckpt.py::EMA (CPU), and the one-launch CUDA implementation against both (GPU, bit-exact)."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "ema_reference.npz")


def make_model():
    """Same construction as oracle/make_golden_ema.py."""
    torch.manual_seed(31)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 7, 3), torch.nn.InstanceNorm2d(7, affine=True),
                              torch.nn.Conv2d(7, 5, 1, bias=False), torch.nn.Linear(9, 4))
    net[2].weight.requires_grad_(False)
    return net


def perturbations(net):
    g = torch.Generator().manual_seed(32)
    return [[torch.randn(p.shape, generator=g) * 0.1 for p in net.parameters()] for _ in range(4)]


def test_oracle_matches_the_reference_fixture():
    from oracle import ema_oracle
    d = np.load(GOLD)
    net = make_model()
    names = [n for n, p in net.named_parameters() if p.requires_grad]
    assert "2.weight" not in names and all(f"s0:{n}" in d for n in names) and "s0:2.weight" not in d
    shadow = {n: p.detach().numpy().copy() for n, p in net.named_parameters() if p.requires_grad}
    for step, deltas in enumerate(perturbations(net)):
        with torch.no_grad():
            for p, dl in zip(net.parameters(), deltas):
                p.add_(dl)
        for n, p in net.named_parameters():
            if p.requires_grad:
                shadow[n] = ema_oracle.ema_update_np(shadow[n], p.detach().numpy(), float(d["decay"]))
                np.testing.assert_array_equal(shadow[n], d[f"s{step}:{n}"], err_msg=f"step {step} {n}")
    for n, p in net.named_parameters():
        np.testing.assert_array_equal(p.detach().numpy(), d[f"final:{n}"])


@pytest.mark.gpu
def test_cuda_ema_is_bit_identical_to_the_reference():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gan_variant_research_b200 as pn
    d = np.load(GOLD)
    net = make_model()
    deltas_all = perturbations(net)
    net = net.cuda()
    ema = pn.EMA(net, decay=float(d["decay"]))
    assert sorted(ema.shadow) == sorted(n for n, p in net.named_parameters() if p.requires_grad)
    for step, deltas in enumerate(deltas_all):
        with torch.no_grad():
            for p, dl in zip(net.parameters(), deltas):
                p.add_(dl.cuda())
        ema.update()
        for n, v in ema.shadow.items():
            np.testing.assert_array_equal(v.cpu().numpy(), d[f"s{step}:{n}"], err_msg=f"step {step} {n}")
    sd = ema.state_dict()
    assert sd["decay"] == float(d["decay"]) and set(sd["shadow"]) == set(ema.shadow)
    trainable = [p for p in net.parameters() if p.requires_grad]      # the frozen parameter is skipped (:19-21)
    v0 = [p._version for p in trainable]
    ema.apply_shadow()
    # the kernels write parameters through raw pointers: the version counters must still move, or anything keyed on
    # them (the encoder-feature cache, feature_reuse.py) would serve values computed with the old weights
    assert all(p._version > a for p, a in zip(trainable, v0))
    for n, p in net.named_parameters():
        np.testing.assert_array_equal(p.detach().cpu().numpy(), d[f"applied:{n}"])
    v1 = [p._version for p in trainable]
    ema.restore()
    assert all(p._version > a for p, a in zip(trainable, v1))
    for n, p in net.named_parameters():
        np.testing.assert_array_equal(p.detach().cpu().numpy(), d[f"final:{n}"])
    with pytest.raises(KeyError):
        ema.restore()
    # checkpoint round trip into a fresh instance
    ema2 = pn.EMA(net, decay=0.5)
    ema2.load_state_dict({"decay": sd["decay"], "shadow": {n: v.cpu() for n, v in sd["shadow"].items()}})
    assert ema2.decay == sd["decay"]
    for n in ema.shadow:
        assert torch.equal(ema2.shadow[n], ema.shadow[n])


@pytest.mark.gpu
def test_cuda_ema_at_generator_size_matches_the_formula():
    """All 48 parameter tensors of the reference generator (11.4 M values; shapes from the fixture): one launch,
    bit-identical to the three-rounding formula evaluated by ATen."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import json
    import gan_variant_research_b200 as pn
    shapes = json.load(open(os.path.join(HERE, "golden", "model_param_shapes.json")))["generator"]
    g = torch.Generator(device="cuda").manual_seed(1)

    class Bag(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes])
    net = Bag()
    assert sum(p.numel() for p in net.parameters()) == 11378179
    ema = pn.EMA(net, decay=0.999)
    want = {n: p.detach().clone() for n, p in net.named_parameters()}
    for _ in range(3):
        with torch.no_grad():
            for p in net.parameters():
                p.mul_(1.01).add_(0.001)
        ema.update()
        for n, p in net.named_parameters():
            want[n] = (1.0 - 0.999) * p.detach() + 0.999 * want[n]
    for n in want:
        assert torch.equal(ema.shadow[n], want[n]), n


def test_cpu_parameters_are_refused():
    import gan_variant_research_b200 as pn
    with pytest.raises(RuntimeError, match="CUDA"):
        pn.EMA(torch.nn.Linear(2, 2))


@pytest.mark.gpu
def test_parameters_moved_to_channels_last_are_refused_loudly():
    """The kernel walks raw storage in the shadow's contiguous order: a model converted to torch.channels_last after
    the EMA was built must raise at the next update(), not average permuted values."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gan_variant_research_b200 as pn
    net = torch.nn.Conv2d(8, 16, 3).cuda()
    ema = pn.EMA(net, decay=0.9)
    ema.update()
    net.to(memory_format=torch.channels_last)
    assert not net.weight.is_contiguous()
    with pytest.raises(RuntimeError, match="contiguous"):
        ema.update()
