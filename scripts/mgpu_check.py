"""torchrun --nproc-per-node N scripts/mgpu_check.py : sharded contraction + distributed sector SVD on
N GPUs must reproduce the single-GPU result bit for bit (contraction) / to 1e-12 (TRG step), and
reports strong-scaling timings of the chi=D contraction."""
import os, sys, time, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
import grassmanntn_b200 as gtn
from grassmanntn_b200 import parallel, _engine as E
dev = torch.device("cuda", local)
g = gtn.gauge2d

def even_block(D, stats, seed):
    gen = torch.Generator(device="cpu"); gen.manual_seed(seed)
    half = D // 2
    b = gtn.zero_block_eo((half,) * 4, (half,) * 4, stats, dtype=complex)
    bt = b._bt
    for p_ in bt.patterns():
        if sum(p_) % 2 == 0:
            v = bt.block_view(p_)
            x = torch.rand(tuple(v.shape) + (2,), generator=gen, dtype=torch.float64)
            v.copy_(torch.view_as_complex(x).to(dev))
        else:
            bt.zero.add(p_)
    return b

out = {}
for D in (32, 64, 128):
    VV, UU = even_block(D, (1, 1, -1, 1), 1), even_block(D, (-1, 1, -1, 1), 2)
    parallel.disable()
    ref = gtn.einsum('lxzk,jzxi->ijkl', VV, UU)
    parallel.enable(min_flops=0.0)
    got = gtn.einsum('lxzk,jzxi->ijkl', VV, UU)
    same = all(torch.equal(ref._bt.block_view(p), got._bt.block_view(p)) for p in ref._bt.off)
    def timeit(n):
        torch.cuda.synchronize(); dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            r = gtn.einsum('lxzk,jzxi->ijkl', VV, UU)
        e.record(); torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    timeit(2)
    ms_sh = timeit(3)
    parallel.disable()
    timeit(1)
    ms_1 = timeit(3)
    out["D%d" % D] = dict(bit_identical=bool(same), ms_single=ms_1, ms_sharded=ms_sh, speedup=ms_1 / ms_sh,
                          TFLOPs_sharded=2.0 * D ** 6 / (ms_sh * 1e-3) / 1e12)
    del VV, UU, ref, got
    torch.cuda.empty_cache()

# full TRG steps: distributed sector SVD + sharded contraction vs single GPU
# two complete chains from the same fixture: replicated tensors must stay bit-identical across the
# ranks, so the sharded chain runs in sharded mode from the start (a tensor produced by an
# un-synchronised local step carries a rank-dependent SVD gauge on its legs)
T0 = g.zcap(g.load_initial_tensor()).toblock()
def chain(sharded):
    if sharded:
        parallel.enable(min_flops=0.0)
    else:
        parallel.disable()
    T, rec = T0, []
    for step in range(4):
        T, n = g.trg(T, 32)
        rec.append((n, complex(g.logZ(T))))
    parallel.disable()
    return rec
ra, rb = chain(False), chain(True)
for step, ((n1, F1), (n2, F2)) in enumerate(zip(ra, rb)):
    out["trg_chi32_step%d" % step] = dict(Tnorm_single=n1, Tnorm_sharded=n2, rel=abs(n1 - n2) / n1,
                                          relF=abs(F1 - F2) / abs(F1))
if rank == 0:
    print(json.dumps(out, indent=1))
dist.destroy_process_group()
