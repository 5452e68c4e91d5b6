import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
import grassmanntn_b200 as gtn
from grassmanntn_b200 import parallel, _ops
g = gtn.gauge2d
T = g.zcap(g.load_initial_tensor()).toblock()
T, _ = g.trg(T, 32); T, _ = g.trg(T, 32)
T1 = gtn.einsum('ijkl->jkli', T); T2 = gtn.einsum('ijkl->klij', T)
parallel.disable()
A = gtn.svd_many([T1, T2], 'ab|cd', 32)
parallel.enable(min_flops=0.0)
B = gtn.svd_many([T1, T2], 'ab|cd', 32)
parallel.disable()
def sv(S): return [np.abs(np.diag(S._bt.block_view(p).cpu().numpy())) for p in ((0,0),(1,1))]
for i in range(2):
    sa, sb = sv(A[i][1]), sv(B[i][1])
    ra = gtn.einsum('abx,xy,ycd->abcd', *A[i]); rb = gtn.einsum('abx,xy,ycd->abcd', *B[i])
    d = (ra + (-1) * rb).norm / ra.norm
    if rank == 0:
        print(i, 'S diff E', np.abs(sa[0]-sb[0]).max(), 'O', np.abs(sa[1]-sb[1]).max(), 'rec diff', d, 'rec vs T', (ra + (-1)*[T1,T2][i]).norm, (rb + (-1)*[T1,T2][i]).norm, flush=True)
        print('   sE', sa[0][12:18], sb[0][12:18])
        print('   U norm', A[i][0].norm, B[i][0].norm, 'V norm', A[i][2].norm, B[i][2].norm)
# per-rank view: is the owner or the receiver wrong?
parallel.enable(min_flops=0.0)
ctxs = [_ops._decompose_prepare(t._bt, 2, "svd") for t in (T1, T2)]
mats = [m for c in ctxs for m in c["mats"]]
usv = _ops._svd_distributed(mats, ctxs, 32, "svd", "block")
parallel.disable()
loc = _ops._svd_local(mats, ctxs, 32, "svd", "block")
for i, (M, (U, s_, Vh), (U2, s2, Vh2)) in enumerate(zip(mats, usv, loc)):
    k = 16
    sd = torch.from_numpy(s_[:k]).to(M.device).to(M.dtype)
    R = (U[:, :k] * sd) @ Vh[:k]
    R2 = (U2[:, :k] * torch.from_numpy(s2[:k]).to(M.device).to(M.dtype)) @ Vh2[:k]
    print('rank', rank, 'prob', i, 'owner', i % world, 'dist rec err', float((R - M).norm() / M.norm()), 'local', float((R2 - M).norm() / M.norm()),
          'U shape', tuple(U.shape), tuple(U2.shape), 'U contiguous', U.is_contiguous(), flush=True)
print(rank, _ops.SVD_PATH_STATS, flush=True)
dist.destroy_process_group()
