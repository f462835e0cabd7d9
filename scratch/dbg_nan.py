import numpy as np, torch, sys
sys.path.insert(0,'.')
import gan_variant_research_b200 as pn
d=np.load('tests/golden/small_nan_image.npz')
n=int(d['n_layers'])
src=[torch.from_numpy(d[f'src{i}']).cuda() for i in range(n)]
tgt=[torch.from_numpy(d[f'tgt{i}']).cuda().requires_grad_() for i in range(n)]
ids=[torch.from_numpy(d[f'ids{i}']).cuda() for i in range(n)]
loss=pn.fused_patchnce(src,tgt,ids,0.07); loss.backward()
g=tgt[0].grad.cpu().numpy(); w=d['grad0']
m=~np.isnan(w)
bad=np.argwhere(((g!=0)!=(w!=0))&m)
print('mismatch idx',bad)
for ix in bad:
    ix=tuple(ix); print(ix,'got',g[ix],'want',w[ix], 'hw', ix[2]*6+ix[3], 'ids', d['ids0'])
