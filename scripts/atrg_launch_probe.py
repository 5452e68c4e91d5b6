"""Per-launch timing of steady-state ATRG steps at chi (single GPU): families and the slowest launch shapes (CUDA events
around every C-ABI launch; the decompositions run on the Python-built plans while the profiler is on)."""
import os, sys, collections, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import _engine as E, _ops
g = gtn.gauge2d
chi = int(os.environ.get("CHI", "128"))
T = g.zcap(g.load_initial_tensor()).toblock()
while tuple(T.effective_shape) != (chi,) * 4:
    T, _ = g.trg(T, chi)
X = T
for i in range(4):
    X = (g.atrg2dy if i % 2 == 0 else g.atrg2dx)(X, X, chi)[0]
torch.cuda.synchronize()
for i in range(2):
    t0 = time.perf_counter()
    E.PROF.start()
    X = (g.atrg2dy if i % 2 == 0 else g.atrg2dx)(X, X, chi)[0]
    E.PROF.stop(detail=True)
    print("step", i, "wall %.1f ms (profiled)" % ((time.perf_counter() - t0) * 1e3), "paths", _ops.SVD_PATH_STATS, "last iters", E.truncated_svd_batch.last_iters)
    fam = collections.defaultdict(list)
    for name, ms, flops, nbytes, desc in E.PROF.detail:
        fam[name].append((ms, flops, desc))
    for name, rows in sorted(fam.items(), key=lambda kv: -sum(r[0] for r in kv[1])):
        print("  %-22s total %8.2f ms in %d launches" % (name, sum(r[0] for r in rows), len(rows)))
        agg = collections.defaultdict(lambda: [0.0, 0, 0])
        for ms, flops, desc in rows:
            a = agg[desc]; a[0] += ms; a[1] += 1; a[2] = flops
        for desc, (ms, n, flops) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:6]:
            print("      %8.3f ms  x%-3d %6.2f TFLOP/s  %s" % (ms, n, (flops * n / (ms * 1e-3) / 1e12) if ms and flops else 0.0, desc))
