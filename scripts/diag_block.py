import sys; sys.path.insert(0,'.'); sys.path.insert(0,'scrna-parameter-estimation_b200'); sys.path.insert(0,'tests')
import numpy as np, torch
from test_gpu_block import _random_matrix
for sizes in ([70,200,129],[300,517,90,1000],[4000, 12000]):
    n_genes=150
    seg, dense, sf, gs = _random_matrix(sizes, n_genes, seed=len(sizes))
    d=seg.device; inv_sf=torch.as_tensor(1.0/sf, device=d); sums=seg.moments(inv_sf)
    idx=np.arange(n_genes)
    got=seg.block_cross(idx, idx, inv_sf, sums).cpu().numpy()
    for r in range(len(sizes)):
        u=dense[gs[r]:gs[r+1]]/sf[gs[r]:gs[r+1],None]; m=u.mean(0,keepdims=True); z=u-m
        n=z.shape[0]
        second=(z**2).sum(0)/n
        e=np.where(second>0, np.round(0.5*np.log2(np.maximum(second,1e-300))), 0.0)
        zs=z*2.0**(-e)
        hi=zs.astype(np.float16).astype(np.float64); lo=((zs-hi)*2048).astype(np.float16).astype(np.float64)
        emu=(hi.T@hi + (hi.T@lo + lo.T@hi)/2048)*np.outer(2.0**e,2.0**e)     # exact accumulation of the 3 products
        want=z.T@z
        scale=np.sqrt(np.outer((z**2).sum(0),(z**2).sum(0)))+1e-300
        e_split=np.abs(emu-want)/scale; e_gpu=np.abs(got[r]-want)/scale; e_acc=np.abs(got[r]-emu)/scale
        i=np.unravel_index(np.argmax(e_gpu), e_gpu.shape)
        print(sizes, 'group',r,'K',n,'split err %.2e  gpu err %.2e (at %s, diag max %.2e, offdiag max %.2e)  accumulation err %.2e'%(e_split.max(), e_gpu.max(), i, np.diag(e_gpu).max(), (e_gpu-np.diag(np.diag(e_gpu))).max(), e_acc.max()))
