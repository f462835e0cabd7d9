"""Find the host stall at the second step after a device sync: time the pieces of steps 0..3 on the host."""
import sys, time, os
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import patchnce as pm
from bench import LAYER_SETS, make_maps, ClockSampler
dev = torch.device('cuda', 0); torch.cuda.set_device(0); B = 64
use_sampler = os.environ.get('SAMPLER', '1') == '1'
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256, [0, 4, 8, 12, 13], math=pn.DEFAULT_MATH)
torch.manual_seed(7)
marks = []
def T(tag): marks.append((tag, time.perf_counter()))
orig_draw = pm.draw_patch_ids_all
def draw(*a, **k):
    T('draw0'); r = orig_draw(*a, **k); T('draw1'); return r
pm.draw_patch_ids_all = draw
orig_apply = pm._FusedPatchNCE.apply
def step(i):
    T(f's{i}')
    for t in tgt: t.grad = None
    T('freed')
    loss = crit(src, tgt)
    T('fwd')
    loss.backward()
    T('bwd')
sampler = ClockSampler(0)
if use_sampler: sampler.start()
for i in range(5): step(-1)
torch.cuda.synchronize()
for rep in range(3):
    marks.clear()
    torch.cuda.synchronize()
    t_sync = time.perf_counter()
    for i in range(6): step(i)
    torch.cuda.synchronize()
    out = []
    prev = t_sync
    for tag, t in marks:
        out.append(f'{tag}+{(t - prev) * 1e6:.0f}')
        prev = t
    print(f'rep {rep} sampler={use_sampler}:', ' '.join(out))
    for _ in range(50): step(-1)
