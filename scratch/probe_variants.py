"""Try descriptor-stride variants if the canonical ones mismatch; prints which (LBO,SBO) match."""
import sys, itertools
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np
from test_umma_selftest_gpu import tile_blob, run_probe
g = torch.Generator().manual_seed(1)
n, k = 256, 32
a = torch.randint(-4, 5, (128, k), generator=g).float(); b = torch.randint(-4, 5, (n, k), generator=g).float()
want = (a @ b.t()).numpy()
for (al, asb), (bl, bsb) in itertools.product([(2048, 128), (128, 2048)], [(4096, 128), (128, 4096)]):
    d, err = run_probe(tile_blob(a), tile_blob(b), (al, asb, 4096), (bl, bsb, 8192), n, k, 0)
    print("K-major A(lbo,sbo)=", (al, asb), "B=", (bl, bsb), "err", err, "match", np.array_equal(d.numpy(), want),
          "maxdiff", float(np.nanmax(np.abs(d.numpy() - want))))
n, k = 32, 256
a = torch.randint(-4, 5, (128, k), generator=g).float(); km = torch.randint(-4, 5, (k, n), generator=g).float()
want = (a @ km).numpy()
for (bl, bsb) in [(128, 4096), (4096, 128)]:
    d, err = run_probe(tile_blob(a), tile_blob(km), (2048, 128, 4096), (bl, bsb, 256), n, k, 1)
    print("MN-major B(lbo,sbo)=", (bl, bsb), "err", err, "match", np.array_equal(d.numpy(), want),
          "maxdiff", float(np.nanmax(np.abs(d.numpy() - want))))
