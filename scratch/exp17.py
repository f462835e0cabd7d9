"""Does the forward (gather) slow down as the box warms up?  One process, repeated timed loops, temperatures."""
import sys, time, subprocess
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from bench import LAYER_SETS, make_maps
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda', 0); B = 64
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256)
def step():
    for t in tgt: t.grad = None
    loss = crit(src, tgt); loss.backward(); return loss
def smi():
    q = 'temperature.gpu,temperature.memory,clocks.sm,clocks.mem,power.draw'
    return subprocess.run(['nvidia-smi', '--query-gpu=' + q, '--format=csv,noheader'], capture_output=True, text=True).stdout.strip()
for _ in range(5): step()
torch.cuda.synchronize()
for rep in range(8):
    if rep == 5:
        print('sleep 20 s'); time.sleep(20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(1500): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 1500
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): step()
        torch.cuda.synchronize()
    rows = {e.key[:40]: e.device_time_total / 5 for e in prof.key_averages() if e.device_time_total > 0 and 'pnce::k_' in e.key}
    print(f'rep {rep}: step {ms*1e3:.1f} us; ' + '  '.join(f'{k.split("::")[-1][:12]}={v:.1f}' for k, v in rows.items()), '|', smi(), flush=True)
