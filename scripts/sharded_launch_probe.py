"""Per-launch timing of one sharded TRG step at chi (torchrun, N ranks): which launches of the replicated / skinny
families cost what (CUDA events around every C-ABI launch, rank 0 prints the slowest ones per family)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
import grassmanntn_b200 as gtn
from grassmanntn_b200 import sharded, _engine as E
g = gtn.gauge2d
chi = int(os.environ.get("CHI", "128"))
T = g.zcap(g.load_initial_tensor()).toblock()
while tuple(T.effective_shape) != (chi,) * 4:
    T, _ = g.trg(T, chi)
sharded.broadcast_tensor(T, 0)
Tl = sharded.shard(T)
del T
for _ in range(4):
    out, _ = sharded.trg(Tl, chi)
torch.cuda.synchronize(); dist.barrier()
E.PROF.start()
out, _ = sharded.trg(Tl, chi)
E.PROF.stop(detail=True)
if rank == 0:
    fam = collections.defaultdict(list)
    for name, ms, flops, nbytes, desc in E.PROF.detail:
        fam[name].append((ms, flops, desc))
    for name, rows in sorted(fam.items(), key=lambda kv: -sum(r[0] for r in kv[1])):
        print("%-22s total %.2f ms in %d launches" % (name, sum(r[0] for r in rows), len(rows)))
        agg = collections.defaultdict(lambda: [0.0, 0, 0])
        for ms, flops, desc in rows:
            a = agg[desc]; a[0] += ms; a[1] += 1; a[2] = flops
        for desc, (ms, n, flops) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:8]:
            print("    %7.3f ms  x%-3d %6.2f TFLOP/s  %s" % (ms, n, (flops * n / (ms * 1e-3) / 1e12) if ms and flops else 0.0, desc))
dist.destroy_process_group()
