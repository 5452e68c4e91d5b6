"""ATRG at large chi, ONE chain sharded over the ranks of a box (BASELINE.json configs[3]: "ATRG 2D Z2 gauge theory
chi=256, contraction sharded across 1/2/4/8 B200").  Run under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/atrg_sharded.py --chi 256

Prologue (untimed): the TRG chain of the Z2 tensor at the same chi until all four legs have reached chi (every rank
runs it, rank 0's tensor is then broadcast: SVD gauges are not bitwise reproducible between processes).  Timed: ATRG
steps alternating x / y (example.py:178-188) on the sharded tensor, CUDA events per step, max over ranks.  With
--check the same steps also run unsharded on every rank (chi <= 128: needs the whole step on one GPU) and Tnorm is
compared.  Prints one JSON line on rank 0."""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--chi", type=int, default=256)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--check", action="store_true")
ap.add_argument("--out", default=None)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
import grassmanntn_b200 as gtn
from grassmanntn_b200 import sharded, _engine as E, _ops
g = gtn.gauge2d
dev = torch.device("cuda", local)
chi = args.chi
T = g.zcap(g.load_initial_tensor()).toblock()
n = 0
while tuple(T.effective_shape) != (chi,) * 4 and n < 8:
    T, _ = g.trg(T, chi)
    n += 1
torch.cuda.synchronize()
sharded.broadcast_tensor(T, 0)
first_x = False                                   # square tensor: example.py's cgxfirst = shape[0] > shape[1] is False -> y first
Tl = sharded.shard(T, 0 if first_x else 1)
Ts = T if args.check else None
if not args.check:
    del T
torch.cuda.empty_cache()
rows = []
st0 = dict(sharded.STATS)
for i in range(args.steps):
    use_x = (i % 2 == 0) == first_x
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    Tl, tn = (sharded.atrg2dx if use_x else sharded.atrg2dy)(Tl, chi)
    e.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([s.elapsed_time(e), wall * 1e3, torch.cuda.max_memory_allocated() / 2 ** 30], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    row = {"step": i, "dir": "x" if use_x else "y", "ms": t[0].item(), "wall_ms": t[1].item(), "Tnorm": tn,
           "peak_mem_GiB_per_rank": t[2].item(), "svd_iterations_last": int(E.truncated_svd_batch.last_iters)}
    if args.check:
        Ts, tn1 = (g.atrg2dx if use_x else g.atrg2dy)(Ts, Ts, chi)
        row["Tnorm_single_gpu"] = tn1
        row["rel_diff"] = abs(tn - tn1) / tn1
    rows.append(row)
    if rank == 0:
        print(json.dumps(row), flush=True)
comm = {k: (sharded.STATS[k] - st0[k]) / max(args.steps, 1) for k in st0}
if rank == 0:
    D = chi
    # algorithmic flops of the three big contractions of a step (non-zero parity sectors): M = C.B, Q = Q1.Q2, T' = H.G
    flops = 2.0 * (D ** 5 * 1 + D ** 6 + D ** 5)
    best = min(r["ms"] for r in rows)
    line = {"workload": "ATRG step (atrg2dx / atrg2dy alternating), 2D Z2 gauge theory, block format, D = chi = %d, "
                        "complex128, site tensor %.1f GiB of even parity blocks" % (chi, chi ** 4 * 8 / 2 ** 30),
            "n_gpus": world, "parallelism": "sharded (grassmanntn_b200/sharded.py)", "steps": rows,
            "best_ms_per_step": best, "steps_per_s": 1e3 / best,
            "contraction_TFLOPs_aggregate": flops / (best * 1e-3) / 1e12,
            "collective_MiB_per_step": {k: v / 2 ** 20 for k, v in comm.items() if k.endswith("bytes")},
            "collectives_per_step": comm["collectives"], "prologue_trg_steps": n}
    print(json.dumps(line))
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(line, open(args.out, "w"), indent=1)
dist.destroy_process_group()
