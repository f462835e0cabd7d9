"""Head GEMM launches with the epilogue switched off piece by piece (experiment build, knob 14): 0 = normal, 1 = epilogue
computes but does not store to global memory, 2 = empty epilogue (MMAs + operand loads only)."""
import sys, ctypes, json
sys.path.insert(0, '.')
import torch
import gan_variant_research_b200 as pn
from gan_variant_research_b200 import _lib
from bench import LAYER_SETS, make_maps, kernel_breakdown
lib = _lib.load(); lib.pnce_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
dev = torch.device('cuda', 0)
src, tgt = make_maps(LAYER_SETS['b5'], 64, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
torch.manual_seed(11)
netF = pn.PatchSampleF(use_mlp=True, nc=256).to(dev); netF.create_mlp(tgt)
def step():
    for t in tgt: t.grad = None
    netF.zero_grad(set_to_none=True)
    loss, _ = pn.patchnce_with_head(netF, src, tgt, 0.07, 256)
    loss.backward()
for mode in (0, 1, 2):
    lib.pnce_debug_set(14, mode)
    for _ in range(3): step()
    print(mode, json.dumps(kernel_breakdown(step)))
lib.pnce_debug_set(14, 0)
