"""PatchNCE path timings at the shapes of every BASELINE.json config (the generator passes of configs
2/3 are the reference's own code and are not part of this path).  One JSON line per case."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_variant_research_b200 as pn

def maps(shapes, b, dtype, seed=1234):
    g = torch.Generator(device='cuda').manual_seed(seed)
    src = [torch.randn(b, *s, device='cuda', generator=g).relu().to(dtype) for s in shapes]
    tgt = [torch.randn(b, *s, device='cuda', generator=g).relu().to(dtype).requires_grad_() for s in shapes]
    return src, tgt

def time_case(name, shapes, b, p, dtype=torch.float32, head=False, math=None, steps=30):
    src, tgt = maps(shapes, b, dtype)
    crit = pn.PatchNCELoss(0.07, p, math=math)
    netF = None
    if head:
        torch.manual_seed(1)
        netF = pn.PatchSampleF(use_mlp=True, nc=256).cuda(); netF.create_mlp(tgt)
    def step():
        for t in tgt: t.grad = None
        if netF is None: loss = crit(src, tgt)
        else:
            netF.zero_grad(set_to_none=True)
            loss, _ = pn.patchnce_with_head(netF, src, tgt, 0.07, p, math=math)
        loss.backward(); return loss
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): l = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    patches = b * sum(min(p, s[1] * s[2]) for s in shapes)
    elem = 4 if dtype == torch.float32 else 2
    alg = b * sum(2 * min(p, s[1] * s[2]) * s[0] * elem + s[0] * s[1] * s[2] * elem for s in shapes)
    print(json.dumps({"case": name, "batch": b, "num_patches": p, "dtype": str(dtype).split('.')[-1], "head": head,
                      "math": math or pn.DEFAULT_MATH, "ms_per_step": round(ms, 4), "patches_per_s": round(patches / ms * 1e3),
                      "hbm_roofline_frac": round(alg / (ms * 1e-3) / 6457.4e9, 4), "loss": round(float(l), 5)}), flush=True)
    del src, tgt; torch.cuda.empty_cache()

def time_case_graph(name, shapes, b, p, steps=200):
    """Same step replayed from a CUDA graph (forward + backward captured once, ids drawn inside the
    graph): what the device needs when the Python host is out of the way."""
    src, tgt = maps(shapes, b, torch.float32)
    crit = pn.PatchNCELoss(0.07, p)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            for t in tgt: t.grad = None
            crit(src, tgt).backward()
    torch.cuda.current_stream().wait_stream(side)
    for t in tgt: t.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss = crit(src, tgt); loss.backward()
    for _ in range(5): graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): graph.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    patches = b * sum(min(p, s[1] * s[2]) for s in shapes)
    alg = b * sum(2 * min(p, s[1] * s[2]) * s[0] * 4 + s[0] * s[1] * s[2] * 4 for s in shapes)
    print(json.dumps({"case": name + " [CUDA graph replay]", "batch": b, "num_patches": p, "dtype": "float32", "head": False,
                      "math": pn.DEFAULT_MATH, "ms_per_step": round(ms, 4), "patches_per_s": round(patches / ms * 1e3),
                      "hbm_roofline_frac": round(alg / (ms * 1e-3) / 6457.4e9, 4), "loss": round(float(loss), 5)}), flush=True)
    del src, tgt, graph; torch.cuda.empty_cache()


R4 = [(64, 256, 256), (256, 64, 64), (256, 64, 64), (128, 128, 128)]
B5 = R4 + [(64, 256, 256)]
B5_512 = [(64, 512, 512), (256, 128, 128), (256, 128, 128), (128, 256, 256), (64, 512, 512)]
time_case("cfg1/2: B=1, 256^2, R4 (what [0,4,8,12,16] returns)", R4, 1, 256, steps=200)
time_case("cfg1/2: B=1, 256^2, B5", B5, 1, 256, steps=200)
time_case("cfg2 AMP: B=1, 256^2, R4, fp16 maps", R4, 1, 256, torch.float16, steps=200)
time_case("cfg3: B=16, 256^2, R4 (FastCUT setting)", R4, 16, 256, steps=100)
time_case("cfg3: B=16, 256^2, B5", B5, 16, 256, steps=100)
time_case("cfg5 per-GPU: B=64, 256^2, B5", B5, 64, 256)
time_case("cfg5 per-GPU AMP: B=64, 256^2, B5, fp16 maps", B5, 64, 256, torch.float16)
time_case("cfg5 per-GPU AMP: B=64, 256^2, B5, bf16 maps, single-pass bf16 MMA", B5, 64, 256, torch.bfloat16, math="tc_bf16")
time_case("cfg5 strong-scaling shard: B=8, 256^2, B5", B5, 8, 256, steps=100)
time_case("north star: B=64, 256^2, B5, netF head nc=256", B5, 64, 256, head=True)
time_case("north star: B=16, 256^2, B5, netF head nc=256", B5, 16, 256, head=True, steps=50)
time_case("cfg4: B=8, 512^2, B5, P=1024 (tcgen05 kernel, key blocks of 256)", B5_512, 8, 1024, steps=5)
time_case("cfg4 shapes at P=256: B=8, 512^2, B5", B5_512, 8, 256, steps=20)
time_case_graph("cfg1/2: B=1, 256^2, B5", B5, 1, 256)
time_case_graph("cfg5 strong-scaling shard: B=8, 256^2, B5", B5, 8, 256)
time_case_graph("cfg3: B=16, 256^2, B5", B5, 16, 256)
time_case_graph("cfg5 per-GPU: B=64, 256^2, B5", B5, 64, 256, steps=50)
