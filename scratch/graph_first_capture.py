"""First CUDA-graph capture of a process after an eager warm-up (scratch/host_floor.py at B = 1 reports
cudaErrorStreamCaptureInvalidated): which call invalidates the capture?"""
import sys, traceback
import torch
sys.path.insert(0, ".")
import gan_variant_research_b200 as pn  # noqa: E402
from gan_variant_research_b200 import patchnce as P  # noqa: E402
from bench import LAYER_SETS, make_maps  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
src, tgt = make_maps(LAYER_SETS["b5"], B, torch.float32, dev, 1234)
tgt = [t.requires_grad_() for t in tgt]
crit = pn.PatchNCELoss(0.07, 256, [0, 4, 8, 12, 13])


def status(tag):
    from cuda.bindings import driver
    st = torch.cuda.current_stream().cuda_stream
    r = driver.cuStreamIsCapturing(st)
    print(f"   [{tag}] capture status {r}", flush=True)


def wrap(mod, name):
    f = getattr(mod, name)

    def g(*a, **k):
        status("before " + name)
        try:
            return f(*a, **k)
        finally:
            status("after  " + name)
    setattr(mod, name, g)


def step():
    for t in tgt:
        t.grad = None
    loss = crit(src, tgt)
    loss.backward()
    return loss


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
if len(sys.argv) > 2:                      # a long eager history first, as scratch/host_floor.py has (n eager + n direct steps)
    nlong = int(sys.argv[2])
    for _ in range(nlong):
        step()
    dt = [t.detach() for t in tgt]
    for _ in range(nlong):
        crit.loss_and_grads(src, dt)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
for name in ("draw_patch_ids_all", "_run_fwd", "_run_bwd", "_shape_plan", "_prepare_maps", "_warn_queue"):
    if hasattr(P, name):
        wrap(P, name)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        step()
    g.replay()
    torch.cuda.synchronize()
    print("first capture OK")
except Exception:
    traceback.print_exc()
    print("first capture FAILED")
g2 = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g2):
        step()
    g2.replay()
    torch.cuda.synchronize()
    print("second capture OK")
except Exception as e:
    print("second capture FAILED", repr(e)[:200])
