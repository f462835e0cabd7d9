"""Golden fixture for the AMP optimiser step from the UNMODIFIED reference:
utils/amp_utils.py::AMPContext.step_optimizer on the Adam of training/sched_optim.py::get_optimizer.
Build container only (needs /root/reference).  There is no GPU here and GradScaler('cuda') disables itself without
one, so the context's scaler attribute is replaced by a CPU GradScaler (same class, same code) -- the method under
test, step_optimizer, runs as written."""
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("PNCE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "GAN_Variant1"))
sys.path.insert(0, REF)
from GAN_Variant1.utils.amp_utils import AMPContext  # noqa: E402
from GAN_Variant1.training.sched_optim import get_optimizer  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# (name, optimiser config, max_grad_norm, scaler on?, per-step gradient magnitudes; 'inf' plants one inf)
SCENARIOS = [
    ("amp_clip", {"lr": 2e-4, "betas": [0.5, 0.999]}, 10.0, True, [0.01, 0.02, 50.0, 0.01, "inf", 0.02, 0.01, 0.03, 0.01]),
    ("noscaler_wd", {"lr": 1e-3, "betas": [0.9, 0.99], "weight_decay": 0.01}, None, False, [0.1, 0.2, 0.1]),
    ("noscaler_clip", {"lr": 2e-4, "betas": [0.5, 0.999]}, 10.0, False, [0.01, 30.0, 0.02]),
]
INIT_SCALE, GROWTH_INTERVAL = 1024.0, 3


def make_model():
    torch.manual_seed(41)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 9, 3), torch.nn.InstanceNorm2d(9, affine=True),
                               torch.nn.Conv2d(9, 5, 1, bias=False), torch.nn.Linear(11, 13))


def gradients(net, mags, seed=42):
    """Per step, one 'true' gradient per parameter (what backward() would leave before loss scaling)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for mag in mags:
        gs = [torch.randn(p.shape, generator=g) * (0.01 if mag == "inf" else mag) for p in net.parameters()]
        if mag == "inf":
            gs[1].view(-1)[3] = float("inf")
        out.append(gs)
    return out


def main():
    out = {}
    for name, cfg, max_norm, use_scaler, mags in SCENARIOS:
        net = make_model()
        opt = get_optimizer(net, cfg)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ctx = AMPContext(enabled=use_scaler)
        if use_scaler:
            ctx.scaler = torch.amp.GradScaler("cpu", init_scale=INIT_SCALE, growth_interval=GROWTH_INTERVAL)
            ctx.scaler.scale(torch.zeros(()))                 # what scale_backward() does first: creates _scale
        for k, gs in enumerate(gradients(net, mags)):
            scale = float(ctx.scaler.get_scale()) if use_scaler else 1.0
            for p, g in zip(net.parameters(), gs):
                p.grad = (g * scale).clone()                   # the scaled gradients backward() leaves
            ctx.step_optimizer(opt, max_grad_norm=max_norm)
            for i, p in enumerate(net.parameters()):
                out[f"{name}:{k}:p{i}"] = p.detach().numpy().copy()
                out[f"{name}:{k}:g{i}"] = p.grad.numpy().copy()
                st = opt.state[p]
                if st:
                    out[f"{name}:{k}:m{i}"] = st["exp_avg"].numpy().copy()
                    out[f"{name}:{k}:v{i}"] = st["exp_avg_sq"].numpy().copy()
                    out[f"{name}:{k}:t{i}"] = np.float32(float(st["step"]))
            if use_scaler:
                out[f"{name}:{k}:scale"] = np.float32(ctx.scaler.get_scale())
                out[f"{name}:{k}:tracker"] = np.int32(int(ctx.scaler._growth_tracker))
    np.savez_compressed(os.path.join(OUT, "amp_step_reference.npz"), **out)
    print("wrote amp_step_reference.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
