"""Zero fill under the loss kernel: sweep of PREFILL_BYTES_PER_ITEM at the bench workload (B=64, B5, fp32)."""
import sys
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import patchnce as pmod
from bench import LAYER_SETS, make_maps
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256)
def step():
    for t in tgt: t.grad = None
    loss = crit(src, tgt); loss.backward(); return loss
ref = None
for rep in range(2):
  for kb in (0, 256, 512, 640, 768, 896, 1024, 1280, 1536, 2048):
    pmod.PREFILL_BYTES_PER_ITEM = kb * 1024
    for _ in range(5): l = step()
    torch.cuda.synchronize()
    n = 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    torch.manual_seed(7); l = step(); gsum = [float(t.grad.double().abs().sum()) for t in tgt]
    if ref is None: ref = (l.item(), gsum)
    assert (l.item(), gsum) == ref, ('results changed', l.item(), gsum, ref)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): step()
        torch.cuda.synchronize()
    rows = {e.key: e.device_time_total / 5 for e in prof.key_averages() if e.device_time_total > 0}
    ks = '  '.join(f'{k.split("::")[-1][:12]}={v:.0f}' for k, v in rows.items() if 'pnce::k_' in k)
    print(f'{kb:5d} KB/item ({kb * 1024 * 640 * B // 64 / 2**30:.2f} GiB): step {ms*1e3:.1f} us  [{ks}]', flush=True)
pmod.PREFILL_BYTES_PER_ITEM = 0
