"""Dense backward variants (debug knob 1: 0 default, 1 fill ceiling, 2 store-first, 4 64-thread CTAs) on plain and on
compressible gradient memory (csrc/comp_alloc.cuh).  Needs the experiment build (PNCE_EXPERIMENTS=1)."""
import sys, ctypes
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from bench import LAYER_SETS, make_maps, kernel_breakdown, timed_steps
lib = _lib.load(); lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layout = sys.argv[2] if len(sys.argv) > 2 else 'nchw'
src, tgt = make_maps(LAYER_SETS['b5'], B, torch.float32, dev, 1234)
if layout == 'nhwc':
    src = [x.contiguous(memory_format=torch.channels_last) for x in src]
    tgt = [x.contiguous(memory_format=torch.channels_last) for x in tgt]
crit = pn.PatchNCELoss(0.07, 256)
def step():
    crit.loss_and_grads(src, tgt)
for comp in (False, True):
    pn.set_gradient_compression(comp)
    for flags in (0, 32):
        lib.pnce_debug_set(1, flags)
        ms = timed_steps(step, 50, 5, 1, dev)
        kb = kernel_breakdown(step, 10)
        dk = {k: v for k, v in kb.items() if 'dense' in k or 'fill' in k or 'scatter' in k}
        print(f'{layout} compressible={comp} dense flags={flags}: step {ms*1e3:.1f} us  {dk}', flush=True)
    lib.pnce_debug_set(1, 0)
