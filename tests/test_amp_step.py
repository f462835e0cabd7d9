"""AMP optimiser step.  CPU: the numpy oracle against the fixture frozen from the
reference's AMPContext.step_optimizer.  GPU: pnce_amp_adam_step against the fixture, against the oracle, and against
torch's own unscale_ / clip_grad_norm_ / GradScaler.step / update sequence on the same device (bit-exact when the clip
does not bind)."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "amp_step_reference.npz")

# must mirror oracle/make_golden_amp_step.py
SCENARIOS = [
    ("amp_clip", {"lr": 2e-4, "betas": (0.5, 0.999)}, 10.0, True, [0.01, 0.02, 50.0, 0.01, "inf", 0.02, 0.01, 0.03, 0.01]),
    ("noscaler_wd", {"lr": 1e-3, "betas": (0.9, 0.99), "weight_decay": 0.01}, None, False, [0.1, 0.2, 0.1]),
    ("noscaler_clip", {"lr": 2e-4, "betas": (0.5, 0.999)}, 10.0, False, [0.01, 30.0, 0.02]),
]
INIT_SCALE, GROWTH_INTERVAL = 1024.0, 3


def make_model():
    torch.manual_seed(41)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 9, 3), torch.nn.InstanceNorm2d(9, affine=True),
                               torch.nn.Conv2d(9, 5, 1, bias=False), torch.nn.Linear(11, 13))


def gradients(net, mags, seed=42):
    g = torch.Generator().manual_seed(seed)
    out = []
    for mag in mags:
        gs = [torch.randn(p.shape, generator=g) * (0.01 if mag == "inf" else mag) for p in net.parameters()]
        if mag == "inf":
            gs[1].view(-1)[3] = float("inf")
        out.append(gs)
    return out


def _close(got, want, rtol=2e-6, what=""):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, what
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan), what
    inf = np.isinf(want)
    assert np.array_equal(got[inf], want[inf]), what
    ok = ~(nan | inf)
    if ok.any():
        scale = np.abs(want[ok]).max()
        assert np.abs(got[ok] - want[ok]).max() <= rtol * scale + 1e-30, (what, np.abs(got[ok] - want[ok]).max(), scale)


@pytest.mark.parametrize("scenario", SCENARIOS, ids=[s[0] for s in SCENARIOS])
def test_oracle_matches_the_reference_fixture(scenario):
    from oracle import amp_step_oracle as orc
    name, cfg, max_norm, use_scaler, mags = scenario
    d = np.load(GOLD)
    net = make_model()
    params = [p.detach().numpy().copy() for p in net.parameters()]
    m = [np.zeros_like(p) for p in params]
    v = [np.zeros_like(p) for p in params]
    steps = [0] * len(params)
    scale, tracker = (np.float32(INIT_SCALE), 0) if use_scaler else (None, 0)
    for k, gs in enumerate(gradients(net, mags)):
        grads = [(g.numpy() * np.float32(scale if use_scaler else 1.0)).astype(np.float32) for g in gs]
        scale, tracker, _ = orc.amp_adam_step_np(params, grads, m, v, steps, scale, tracker, lr=cfg["lr"],
                                                 betas=cfg["betas"], weight_decay=cfg.get("weight_decay", 0.0),
                                                 max_grad_norm=max_norm, growth_interval=GROWTH_INTERVAL)
        for i in range(len(params)):
            _close(params[i], d[f"{name}:{k}:p{i}"], what=f"{name} step {k} param {i}")
            _close(grads[i], d[f"{name}:{k}:g{i}"], what=f"{name} step {k} grad {i}")
            _close(m[i], d[f"{name}:{k}:m{i}"], what=f"{name} step {k} exp_avg {i}")
            _close(v[i], d[f"{name}:{k}:v{i}"], what=f"{name} step {k} exp_avg_sq {i}")
            assert steps[i] == float(d[f"{name}:{k}:t{i}"])
        if use_scaler:
            assert float(scale) == float(d[f"{name}:{k}:scale"]) and tracker == int(d[f"{name}:{k}:tracker"])


def test_constructor_rejects_what_it_cannot_mirror():
    from gan_variant_research_b200 import FusedAdamStep
    net = make_model()
    with pytest.raises(TypeError):
        FusedAdamStep(torch.optim.SGD(net.parameters(), lr=0.1))
    with pytest.raises(NotImplementedError):
        FusedAdamStep(torch.optim.Adam(net.parameters(), amsgrad=True))
    with pytest.raises(NotImplementedError):
        FusedAdamStep(torch.optim.Adam([{"params": net[0].parameters()}, {"params": net[3].parameters()}]))
    opt = torch.optim.Adam(net.parameters())
    st = FusedAdamStep(opt)
    for p in net.parameters():
        p.grad = torch.zeros_like(p)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        st.step()                                   # CPU parameters: fails loudly


# ------------------------------------------------------------------------------------------------------------------
def _gpu_run(scenario, mode):
    """mode 'ours': FusedAdamStep;  'torch': the reference's step_optimizer body on stock torch."""
    import gan_variant_research_b200 as pn
    name, cfg, max_norm, use_scaler, mags = scenario
    net = make_model().cuda()
    opt = torch.optim.Adam(net.parameters(), **cfg)
    scaler = torch.amp.GradScaler("cuda", init_scale=INIT_SCALE, growth_interval=GROWTH_INTERVAL, enabled=use_scaler)
    if use_scaler:
        scaler.scale(torch.zeros((), device="cuda"))
    stepper = pn.FusedAdamStep(opt, scaler if use_scaler else None, max_norm) if mode == "ours" else None
    trace = []
    params = list(net.parameters())
    for k, gs in enumerate(gradients(net, mags)):
        scale = scaler.get_scale() if use_scaler else 1.0
        for p, g in zip(params, gs):
            p.grad = (g * scale).cuda()
        if mode == "ours":
            stepper.step()
        else:                                                            # amp_utils.py:29-41
            if max_norm is not None:
                scaler.unscale_(opt)
                torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], max_norm)
            scaler.step(opt)
            scaler.update()
        rec = {"p": [p.detach().cpu().numpy().copy() for p in params],
               "g": [p.grad.cpu().numpy().copy() for p in params],
               "m": [opt.state[p]["exp_avg"].cpu().numpy().copy() for p in params],
               "v": [opt.state[p]["exp_avg_sq"].cpu().numpy().copy() for p in params],
               "t": [float(opt.state[p]["step"]) for p in params],
               "scale": scaler.get_scale() if use_scaler else None,
               "tracker": int(scaler._growth_tracker) if use_scaler else None}
        trace.append(rec)
    return trace


@pytest.mark.gpu
@pytest.mark.parametrize("scenario", SCENARIOS, ids=[s[0] for s in SCENARIOS])
def test_cuda_step_matches_fixture_and_torch(scenario):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    name, cfg, max_norm, use_scaler, mags = scenario
    d = np.load(GOLD)
    ours, ref = _gpu_run(scenario, "ours"), _gpu_run(scenario, "torch")
    clipped_before = False
    for k, (a, b) in enumerate(zip(ours, ref)):
        n = len(a["p"])
        clipped_now = mags[k] not in ("inf",) and mags[k] >= 30.0
        for i in range(n):
            # against the reference fixture (CPU torch): fp32 rounding
            for key, tag in (("p", "p"), ("g", "g"), ("m", "m"), ("v", "v")):
                _close(a[key][i], d[f"{name}:{k}:{tag}{i}"], rtol=3e-6, what=f"{name} step {k} {key}{i} vs fixture")
            assert a["t"][i] == float(d[f"{name}:{k}:t{i}"]) == b["t"][i]
            # against torch on the same GPU: identical bits until a clip has bound (its norm is summed in another order)
            for key in ("p", "g", "m", "v"):
                if clipped_now or clipped_before:
                    _close(a[key][i], b[key][i], rtol=3e-6, what=f"{name} step {k} {key}{i} vs torch")
                else:
                    assert np.array_equal(a[key][i], b[key][i], equal_nan=True), f"{name} step {k} {key}{i} not bit-identical"
        clipped_before = clipped_before or clipped_now
        if use_scaler:
            assert a["scale"] == b["scale"] == float(d[f"{name}:{k}:scale"])
            assert a["tracker"] == b["tracker"] == int(d[f"{name}:{k}:tracker"])


@pytest.mark.gpu
def test_state_dict_round_trip_and_plain_torch_step_afterwards():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gan_variant_research_b200 as pn
    net = make_model().cuda()
    opt = torch.optim.Adam(net.parameters(), lr=2e-4, betas=(0.5, 0.999))
    st = pn.FusedAdamStep(opt, None, 10.0)
    gs = gradients(net, [0.01, 0.02, 0.01])
    for p, g in zip(net.parameters(), gs[0]):
        p.grad = g.cuda()
    st.step()
    assert float(st.last_total_norm()) == pytest.approx(
        float(torch.sqrt(sum((g.double() ** 2).sum() for g in gs[0]))), rel=1e-6)
    import copy
    sd = copy.deepcopy(opt.state_dict())                        # what utils/io_ckpt.py saves (load_state_dict aliases tensors)
    assert all(float(s["step"]) == 1.0 for s in sd["state"].values())
    net2 = make_model().cuda()
    net2.load_state_dict(net.state_dict())
    opt2 = torch.optim.Adam(net2.parameters(), lr=2e-4, betas=(0.5, 0.999))
    opt2.load_state_dict(sd)
    st2 = pn.FusedAdamStep(opt2, None, 10.0)
    for (p, p2), g in zip(zip(net.parameters(), net2.parameters()), gs[1]):
        p.grad = g.cuda(); p2.grad = g.cuda()
    st.step(); st2.step()
    for p, p2 in zip(net.parameters(), net2.parameters()):
        assert torch.equal(p, p2)
    for p, g in zip(net.parameters(), gs[2]):
        p.grad = g.cuda()
    opt.step()                                                  # torch's own (capturable) step still runs on this state
    assert all(float(opt.state[p]["step"]) == 3.0 for p in net.parameters())


@pytest.mark.gpu
def test_amp_step_optimizer_is_a_method_replacement():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gan_variant_research_b200 as pn

    class Ctx:                                                  # the two attributes AMPContext has (amp_utils.py:8-10)
        def __init__(self):
            self.enabled = True
            self.scaler = torch.amp.GradScaler("cuda", init_scale=8.0, growth_interval=2)
        step_optimizer = pn.amp_step_optimizer

    net = make_model().cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    ctx = Ctx()
    before = [p.detach().clone() for p in net.parameters()]
    for k in range(2):
        loss = sum((p ** 2).sum() for p in net.parameters())
        opt.zero_grad()
        ctx.scaler.scale(loss).backward()
        ctx.step_optimizer(opt, max_grad_norm=10.0)
    assert ctx.scaler.get_scale() == 16.0                       # two clean steps with growth_interval=2
    assert all(not torch.equal(a, p) for a, p in zip(before, net.parameters()) if a.dim() > 1)   # zero biases stay zero
