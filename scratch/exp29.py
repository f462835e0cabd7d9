"""AMP optimiser step on a ResNet-9-sized parameter set: the reference's step_optimizer body on stock torch vs FusedAdamStep."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import gan_variant_research_b200 as pn
from standin_generator import StandInGenerator
def setup():
    torch.manual_seed(0)
    gen = StandInGenerator(ngf=64, n_blocks=9).cuda()
    opt = torch.optim.Adam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999))
    # scale 1 and a growth interval that is never reached: the gradients stay valid from call to call (unscale by 1,
    # clip coefficient 1), so nothing but the step itself sits in the timed loop
    sc = torch.amp.GradScaler('cuda', init_scale=1.0, growth_interval=10 ** 9); sc.scale(torch.zeros((), device='cuda'))
    for p in gen.parameters(): p.grad = torch.randn_like(p) * 1e-4
    return gen, opt, sc
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t_host = (time.perf_counter() - t0) / n; torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, t_host * 1e6
gen, opt, sc = setup()
params = [p for p in gen.parameters()]
print('tensors', len(params), 'parameters', sum(p.numel() for p in params))
def ref():
    sc.unscale_(opt); torch.nn.utils.clip_grad_norm_(params, 10.0); sc.step(opt); sc.update()
gen2, opt2, sc2 = setup()
st = pn.FusedAdamStep(opt2, sc2, 10.0)
def ours():
    st.step()
r = timed(ref); o = timed(ours)
print(f'reference sequence on stock torch: {r[0]:.0f} us per call on the device, {r[1]:.0f} us of host time per call')
print(f'FusedAdamStep                    : {o[0]:.0f} us per call on the device, {o[1]:.0f} us of host time per call')
for p, p2 in zip(gen.parameters(), gen2.parameters()):
    assert torch.equal(p, p2), 'parameters diverged'
print('parameters bit-identical after', 55, 'steps each')
from torch.profiler import profile, ProfilerActivity
for name, fn in (('reference', ref), ('ours', ours)):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn(); torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type.name == 'CUDA']
    print(name, 'GPU launches', len(ev), 'kernel time %.0f us' % sum(e.device_time for e in ev))
