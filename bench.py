#!/usr/bin/env python
"""bench.py -- TRG coarse-graining steps/sec of the 2D Z2 gauge theory at chi (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--chi 128] [--impl ours|reference]

One "step" = one Levin-Nave TRG coarse-graining step (reference gauge2d_block.trg, gauge2d_block.py:1649-1755) of the
chi-saturated site tensor (chi^4 complex128, block format) of the Z2 (K=2), N_f=1, beta=m=q=a=1, mu=0 model.  The
default chi = 128 is the regime BASELINE.json's north_star names (chi >= 128: 4 GiB site tensor, 8192 x 8192 sector
matrices); --chi 32 / 64 run the smaller configurations (chi = 32 = example_block.py; its numbers also appear in
`extra` of the default run).  Prints ONE JSON line.

  value      steps/s with the tensor already resident in HBM (CUDA events per step, L2 flushed between steps, max
             over ranks)
  e2e        steps/s through the public API from HOST buffers: pinned host tensor in the block storage layout
             (checkpoint.HostTensor) -> H2D -> gauge2d.trg -> D2H of T' and Tnorm, every step
  roofline   the kernel family with the largest share of the step (CUDA events around every C-ABI launch in a second
             pass over the same steps) against its bound: FP64 tensor pipe (yard-stick: cuBLAS ZGEMM timed in this
             run; nominal 40 TFLOP/s stated beside it) or HBM (MEASURED_PEAKS.json)
  cpu_baseline  the numpy oracle port (oracle/gtn_oracle.py trg_block) on the host cores, one bounded sample
             (rank 0, N=1): full steps for chi <= 64; for chi = 128 the same estimator as --impl reference (a quarter
             sample of the D = chi = 64 step, below), two samples

--impl reference runs ONLY the CPU oracle port (the reference is Python and does not travel to the GPU box; see
DESIGN.md) with all host threads on the same config/metric; at chi = 128 every "step" is a bounded quarter sample
of the D = 64 step (one 2048^2 sector SVD + two 2048^3 sector GEMMs), scaled to the full chi = 128 step.
N > 1 (torchrun): ONE chain sharded over the N GPUs (gauge2d.trg on a leg-sharded tensor: grassmanntn_b200/sharded.py),
value = K / max-over-ranks time, scaling "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np

CPU_LEVEL_MAX = 64          # the CPU oracle port runs full steps up to this D = chi


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chi", type=int, default=128)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent replicas instead of one sharded chain")
    return ap.parse_args()


def load_z2():
    path = os.path.join(ROOT, "tests", "golden", "z2_initial_tensor.npz")
    if os.path.exists(path):
        z = np.load(path)
        return z["data"], tuple(int(s) for s in z["statistics"]), "Z2 gauge initial tensor (reference fixture)"
    # fixture not generated yet: a random Grassmann-even tensor of the same shape
    import gtn_oracle as O
    rng = np.random.RandomState(0)
    T = O.random_dense((8, 8, 8, 8, 2, 2), (1, 1, -1, -1, 0, 0), dtype=complex, rng=rng)
    return T.data, T.statistics, "random Grassmann-even tensor (Z2 fixture missing)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        self.stop_flag = True
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
#  CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def best_blas_threads():
    """OpenBLAS/LAPACK gesdd can be pathologically slow with all threads on a busy or CPU-limited
    host (seen: 17 s vs 0.24 s for one 512x512 complex SVD).  Give the CPU baseline its best
    setting: time one SVD with 1 thread and with all threads and keep the faster."""
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        return None, os.cpu_count()
    a = np.random.RandomState(0).rand(256, 256) + 1j
    best = None
    for n in (1, os.cpu_count()):
        with threadpool_limits(limits=n):
            np.linalg.svd(a)
            t0 = time.perf_counter()
            np.linalg.svd(a)
            dt = time.perf_counter() - t0
        if best is None or dt < best[1]:
            best = (n, dt)
    return threadpool_limits, best[0]


def cpu_level(chi):
    """(D = chi level the CPU port runs at, factor to the full step): every O(D^6) term scales by (chi/level)^6"""
    if chi <= CPU_LEVEL_MAX:
        return chi, 1.0
    return CPU_LEVEL_MAX, (chi / CPU_LEVEL_MAX) ** 6


def cpu_saturated(O, data, stats, level):
    """the first site tensor of the chain with all four legs at `level` (same rule as the GPU arm)"""
    B = O.Blocks.from_dense(O.zcap(O.Dense(data, stats)))
    for _ in range(8):
        if tuple(B.effective_shape) == (level,) * 4:
            break
        B, _ = O.trg_block(B, level)
    return B


def cpu_quarter_setup(O, B):
    """the even-sector matrix of the (ab|cd) matricisation of T1 = T[jkli] (magnitudes only: a timing sample)"""
    T1 = O.einsum_block("ijkl->jkli", B)

    def mat(p):
        b = T1.blocks[p]
        return b.reshape(b.shape[0] * b.shape[1], -1)
    return np.block([[mat((0, 0, 0, 0)), mat((0, 0, 1, 1))], [mat((1, 1, 0, 0)), mat((1, 1, 1, 1))]])


def cpu_reference_run(args, data, stats, quick=False):
    """-> dict(value, ms, shape, threads, sample).  quick: the N=1 GPU arm's cpu_baseline (one bounded sample)."""
    import gtn_oracle as O
    limiter, nthreads = best_blas_threads()
    level, scale = cpu_level(args.chi)

    def body():
        B = cpu_saturated(O, data, stats, level)
        steps, warm = (3, 1) if quick else (args.steps, args.warmup)
        if scale == 1.0:
            for _ in range(warm):
                O.trg_block(B, level)
            t0 = time.perf_counter()
            for _ in range(steps):
                O.trg_block(B, level)
            dt = (time.perf_counter() - t0) / steps
            return dt, "%d full TRG steps at chi=%d (oracle/gtn_oracle.py trg_block; numpy/LAPACK)" % (steps, level)
        if quick:
            steps, warm = 2, 1          # the same estimator as the reference arm (below), two samples
        M = cpu_quarter_setup(O, B)
        for _ in range(min(warm, 1)):
            np.linalg.svd(M, full_matrices=False)
        t0 = time.perf_counter()
        for _ in range(steps):
            np.linalg.svd(M, full_matrices=False)
            M @ M
            M @ M
        dt = (time.perf_counter() - t0) / steps
        return dt * 4 * scale, ("per step a QUARTER sample of the D=chi=%d step: LAPACK SVD of one %dx%d sector matrix of "
                                "T1 + two of its eight %d^3 sector GEMMs (%.2f s), x 4 x (chi/%d)^6 = %g (the smaller "
                                "O(D^5) terms of a step are left out: favours the CPU)"
                                % (level, M.shape[0], M.shape[1], M.shape[0], dt, level, 4 * scale))
    if limiter is not None:
        with limiter(limits=nthreads):
            t_step, sample = body()
    else:
        t_step, sample = body()
    return {"value": 1.0 / t_step, "ms": t_step * 1e3, "threads": nthreads, "sample": sample}


def settle(g, fn, max_steps=12, quiet_needed=2):
    """Run fn() until `quiet_needed` consecutive calls replayed a recorded step graph without recording one -- or
    max_steps calls, for workloads that never use a step graph.  A graph dropped for a re-derivation of the SVD
    iteration counts is recorded again within a few steps; timing loops start after that.  Returns the calls made."""
    st, quiet, n = g.STEP_GRAPH_STATS, 0, 0
    while n < max_steps and quiet < quiet_needed:
        c0, r0 = st["captured"], st["replayed"]
        fn()
        n += 1
        quiet = quiet + 1 if (st["captured"] == c0 and st["replayed"] > r0) else 0
    return n


def config_dict(args, label, parallelism):
    chi = args.chi
    return {"workload": "TRG step (gauge2d_block.trg), 2D Z2 gauge theory K=2 Nf=1 beta=m=q=a=1 mu=0, block format, "
                        "chi=%d, site tensor %s complex128 (%.2f GiB of even parity blocks)"
                        % (chi, "x".join([str(chi)] * 4), chi ** 4 * 16 / 2 / 2 ** 30),
            "input": label, "chi": chi, "parallelism": parallelism,
            "l2": "flushed between timed steps (256 MiB write)",
            "batch": "every timed step processes the same site tensor (the first of the chain with all legs at chi), "
                     "in both arms"}


def saturate(g, T, chi):
    """the first tensor of the TRG chain whose four legs have reached chi (untimed prologue)"""
    n = 0
    while tuple(T.effective_shape) != (chi,) * 4 and n < 8:
        T, _ = g.trg(T, chi)
        n += 1
    return T, n


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    data, stats, label = load_z2()

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args, data, stats)
        cores = "%d BLAS threads (best of 1/all) on %d logical cores" % (r["threads"], os.cpu_count())
        line = {"impl": "reference", "metric": "TRG coarse-grain steps/sec at chi", "value": r["value"], "unit": "steps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms"],
                "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
                "dtype": "complex128", "data": "synthetic",
                "config": config_dict(args, label, "1 CPU process, all host threads"),
                "cpu_baseline": {"value": r["value"], "unit": "steps/s", "cores": cores, "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    if world > 1 and not args.replicas:
        return main_sharded(args, data, stats, label)
    return main_single(args, data, stats, label)


def main_single(args, data, stats, label):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    import grassmanntn_b200 as gtn
    from grassmanntn_b200 import _engine as E, checkpoint as ck
    g = gtn.gauge2d
    dev = torch.device("cuda", local)
    chi = args.chi

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    T, sat_steps = saturate(g, g.zcap(gtn.dense(data, statistics=stats)).toblock(), chi)
    # untimed prologue, continued: reach the engine's steady state on this layout (iteration counts of the truncated SVD
    # settled, its schedules -- and at chi <= 64 the whole-step graph -- recorded): one-off set-up, not a step.
    # Every timed step works on the SAME input T, like the CPU arm does.
    for _ in range(24 if chi <= 64 else 6):
        g.trg(T, chi)
    settle(g, lambda: g.trg(T, chi), max_steps=12 if chi <= 64 else 2)
    g.freeze(True)                           # timing: learnt iteration counts and recorded graphs stay as they are
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # ---- clocks / throttle reasons under load: sampled on rank 0 only, from the warm-up to the end of the e2e loop
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    # ---- device-resident steps
    for _ in range(args.warmup):
        flush.fill_(1)
        g.trg(T, chi)
    barrier()
    n0 = gtn.launch_count()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        X, Tn = g.trg(T, chi)
        e.record()
        evs.append((s, e))
    barrier()
    launches = gtn.launch_count() - n0
    t_dev = sum(s.elapsed_time(e) for s, e in evs) * 1e-3
    del X
    # ---- per-kernel-family shares: the same steps once more with CUDA events around every launch (the timed loop
    #      replays the truncated-SVD schedule as a CUDA graph, where single launches cannot be bracketed; with the
    #      profiler on the engine launches the same kernels one by one).  At chi >= 64 a launch lasts 0.1-250 ms, so
    #      the event overhead (~10 us) does not distort the shares; at chi = 32 see DESIGN.md.
    nprof = min(args.steps, 5 if chi > 64 else args.steps)
    E.PROF.start()
    for _ in range(nprof):
        flush.fill_(1)
        g.trg(T, chi)
    prof = E.PROF.stop()
    g.freeze(False)

    # ---- end to end from host buffers (block storage layout in pinned memory, one copy each way per step)
    h_in = ck.to_host(T)
    torch.cuda.synchronize()
    box = [None]

    def e2e_step():
        Xb = ck.from_host(h_in)
        Y, Tn_ = g.trg(Xb, chi)
        box[0] = ck.to_host(Y, out=box[0])
        torch.cuda.current_stream().synchronize()
        return Tn_
    for _ in range(3):
        e2e_step()
    settle(g, e2e_step, max_steps=12 if chi <= 64 else 1)
    g.freeze(True)
    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e_serial = time.perf_counter() - t0
    # the same K steps through the streaming API (checkpoint.stream_steps): every step still copies ITS input from
    # pinned host memory and ITS result back, but the copies of neighbouring steps overlap the kernels
    ck.stream_steps([h_in] * max(args.warmup, 2), lambda X: g.trg(X, chi))
    barrier()
    t0 = time.perf_counter()
    res_ = ck.stream_steps([h_in] * args.steps, lambda X: g.trg(X, chi))
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    box[0] = res_[-1][0]
    if t_e2e_serial < t_e2e:
        t_e2e = t_e2e_serial
    g.freeze(False)
    clocks = sampler.result() if sampler is not None else None
    h2d, d2h = h_in.nbytes, box[0].nbytes + 8
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30

    tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = tt.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    T_atrg = T if (not args.no_micro and chi >= 128) else None
    del T, h_in, box
    torch.cuda.empty_cache()

    hbm_peak, peak_src = peaks()
    f64_peak, f64_src = fp64_tensor_peak(torch, dev)
    roofline, shares, step_frac = roofline_of(prof, nprof, t_dev / args.steps, chi, hbm_peak, peak_src, f64_peak, f64_src)
    from grassmanntn_b200 import _ops
    extra = {"step_graph": dict(g.STEP_GRAPH_STATS), "speculation": dict(g.SPEC_STATS), "kernel_shares": shares,
             "whole_step_vs_fp64_yardstick": step_frac, "saturation_steps": sat_steps, "peak_mem_GiB": peak_mem,
             "jacobi_sweeps_last": E.batched_svd.last_sweeps, "svd_paths": dict(_ops.SVD_PATH_STATS),
             "trunc_refinements_last": E.truncated_svd_batch.last_iters}
    if not args.no_micro and chi >= 128:
        # the reference's default algorithm (ATRG, example.py:178-188) at the same size: alternating y / x steps from the
        # same saturated tensor, time of the later steps (the first ones derive the iteration counts)
        try:
            # a MOVING chain with adaptation on (the headline repeats one tensor): four consecutive TRG steps
            X, ts = T_atrg, []
            for i in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                X = g.trg(X, chi)[0]
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            extra["trg_block_moving_chain_chi%d_ms" % chi] = ts
        except Exception as ex:
            extra["trg_block_moving_chain_chi%d_ms" % chi] = {"error": repr(ex)[:200]}
        try:
            X, ts = T_atrg, []
            for i in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                X = (g.atrg2dy if i % 2 == 0 else g.atrg2dx)(X, X, chi)[0]
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            extra["atrg_block_chi%d_ms_per_step" % chi] = {"steps": ts, "best": min(ts[2:])}
            del X, T_atrg
            torch.cuda.empty_cache()
        except Exception as ex:
            extra["atrg_block_chi%d_ms_per_step" % chi] = {"error": repr(ex)[:200]}
    if not args.no_micro:
        extra["microbench"] = mb = microbench(gtn, E, torch, dev, args, hbm_peak)
        extra["other_workloads"] = other_workloads(gtn, torch, data, stats, 32)
        extra["rooflines_at_scale"] = {
            "sign_permute_D128": {"bound": "hbm", "achieved": mb["sign_permute_D128"]["GBps"], "peak": hbm_peak,
                                  "unit": "GB/s", "frac": mb["sign_permute_D128"]["GBps"] / hbm_peak,
                                  "peak_source": peak_src, "algorithmic_bytes": mb["sign_permute_D128"]["bytes"],
                                  "traffic": 8546507000,
                                  "traffic_source": "profiles/r1_sign_permute_D128_ncu_full.txt (dram read + write per launch)"}}

    cpu = None
    if world == 1 and not args.no_cpu:
        r = cpu_reference_run(args, data, stats, quick=True)
        cpu = {"value": r["value"], "unit": "steps/s",
               "cores": "%d BLAS threads (best of 1/all) on %d logical cores" % (r["threads"], os.cpu_count()),
               "kind": "port", "sample": r["sample"]}

    line = {"metric": "TRG coarse-grain steps/sec at chi", "value": world * args.steps / t_dev, "unit": "steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128",
            "data": "synthetic", "config": config_dict(args, label, "replicas x%d" % world), "clocks": clocks,
            "e2e": {"value": world * args.steps / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "serial_value": world * args.steps / t_e2e_serial,
                    "mode": "streamed through checkpoint.stream_steps (every step copies its input from pinned host "
                            "memory and its result back; the copies of neighbouring steps overlap the kernels); "
                            "serial_value = the same steps with copy-in, step and copy-out strictly one after the other"},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "extra": extra}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main_sharded(args, data, stats, label):
    """N > 1: ONE TRG chain, site tensor sharded along its first leg over the N ranks (grassmanntn_b200/sharded.py):
    column-sharded truncated SVD (all-reduce of sketch panels), all-gather of the isometries, output-row sharded
    contraction.  Same step, same input tensor as the N = 1 line: strong scaling."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    import grassmanntn_b200 as gtn
    from grassmanntn_b200 import _engine as E, checkpoint as ck, sharded
    g = gtn.gauge2d
    dev = torch.device("cuda", local)
    chi = args.chi

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    # prologue: every rank runs the short chain to the first chi^4 tensor, then rank 0's copy becomes everybody's
    # (SVD gauges are not bitwise reproducible between processes)
    T, sat_steps = saturate(g, g.zcap(gtn.dense(data, statistics=stats)).toblock(), chi)
    sharded.broadcast_tensor(T, 0)
    Tl = sharded.shard(T)
    # one single-GPU step of the same tensor on rank 0's data (every rank: it is the parity reference of the line)
    _, tn_single = g.trg(T, chi)
    del T
    torch.cuda.empty_cache()
    for _ in range(6):                        # iteration counts of the sharded decomposition settle
        sharded.trg(Tl, chi)
    g.freeze(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    for _ in range(args.warmup):
        flush.fill_(1)
        sharded.trg(Tl, chi)
    barrier()
    n0 = gtn.launch_count()
    st0 = dict(sharded.STATS)
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        X, tn_sharded = sharded.trg(Tl, chi)
        e.record()
        evs.append((s, e))
    barrier()
    launches = gtn.launch_count() - n0
    comm = {k: (sharded.STATS[k] - st0[k]) / args.steps for k in st0}
    t_dev = sum(s.elapsed_time(e) for s, e in evs) * 1e-3
    del X
    nprof = min(args.steps, 5)
    E.PROF.start()
    for _ in range(nprof):
        flush.fill_(1)
        sharded.trg(Tl, chi)
    prof = E.PROF.stop()
    # ---- end to end: every rank streams ITS shard from / to pinned host memory
    h_in = ck.to_host(Tl)
    torch.cuda.synchronize()
    box = [None]

    def e2e_step():
        Xb = ck.from_host(h_in)
        Xb._shard_full = Tl._shard_full
        Y, tn = sharded.trg(Xb, chi)
        box[0] = ck.to_host(Y, out=box[0])
        torch.cuda.current_stream().synchronize()
    for _ in range(max(args.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t_e2e_serial = time.perf_counter() - t0

    def prep(X):
        X._shard_full = Tl._shard_full
    ck.stream_steps([h_in] * max(args.warmup, 2), lambda X: sharded.trg(X, chi), prepare=prep)
    barrier()
    t0 = time.perf_counter()
    res_ = ck.stream_steps([h_in] * args.steps, lambda X: sharded.trg(X, chi), prepare=prep)
    barrier()
    t_e2e = min(time.perf_counter() - t0, t_e2e_serial)
    box[0] = res_[-1][0]
    g.freeze(False)
    clocks = sampler.result() if sampler is not None else None
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
    tt = torch.tensor([t_dev, t_e2e, peak_mem], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e, peak_mem = tt.tolist()
    h2d, d2h = h_in.nbytes * world, (box[0].nbytes + 8) * world
    if rank != 0:
        dist.destroy_process_group()
        return
    hbm_peak, peak_src = peaks()
    f64_peak, f64_src = fp64_tensor_peak(torch, dev)
    roofline, shares, step_frac = roofline_of(prof, nprof, t_dev / args.steps, chi, hbm_peak, peak_src, f64_peak, f64_src)
    step_frac["note"] += "; rank 0's share of the flops over the step time (x N for the job)"
    coll = {k: v for k, v in shares.items() if k.startswith("nccl")}
    limiting = max(coll, key=lambda k: coll[k]["ms_per_step"]) if coll else None
    line = {"metric": "TRG coarse-grain steps/sec at chi", "value": args.steps / t_dev, "unit": "steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128",
            "data": "synthetic",
            "config": config_dict(args, label, "sharded: site tensor split along its first leg over %d ranks "
                                               "(column-sharded truncated SVD, isometry all-gather, output-row "
                                               "sharded contraction)" % world),
            "clocks": clocks,
            "e2e": {"value": args.steps / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "serial_value": args.steps / t_e2e_serial,
                    "mode": "streamed through checkpoint.stream_steps (per rank: its shard in, its shard of the result out; "
                            "copies of neighbouring steps overlap the kernels)"},
            "gpu_launches": launches,
            "sharded": {"ranks": world, "tnorm_rel_diff_vs_single_gpu": abs(tn_sharded - tn_single) / tn_single,
                        "allreduce_MiB_per_step": comm["allreduce_bytes"] / 2 ** 20,
                        "allgather_MiB_per_step": comm["allgather_bytes"] / 2 ** 20,
                        "broadcast_MiB_per_step": comm["broadcast_bytes"] / 2 ** 20,
                        "collectives_per_step": comm["collectives"],
                        "collective_ms_per_step": {k: v["ms_per_step"] for k, v in coll.items()},
                        "limiting_collective": limiting, "peak_mem_GiB_per_rank": peak_mem,
                        "svd_iterations": E.truncated_svd_batch.last_iters},
            "roofline": roofline, "cpu_baseline": None,
            "extra": {"kernel_shares": shares, "whole_step_vs_fp64_yardstick": step_frac, "saturation_steps": sat_steps}}
    print(json.dumps(line))
    dist.destroy_process_group()


def roofline_of(prof, nprof, step_s, chi, hbm_peak, peak_src, f64_peak, f64_src):
    """(roofline of the dominant kernel family, per-family shares, whole-step fraction of the FP64 yard-stick)"""
    tot_ms = sum(v["ms"] for v in prof.values())
    shares = {k: {"ms_per_step": v["ms"] / nprof, "launches_per_step": v["launches"] / nprof,
                  "share": v["ms"] / tot_ms if tot_ms else None,
                  "algorithmic_GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] and v["bytes"] else None,
                  "algorithmic_TFLOPs": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] and v["flops"] else None}
              for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    dom = max(prof, key=lambda k: prof[k]["ms"])
    d = prof[dom]
    per_launch_ms = d["ms"] / max(d["launches"], 1)
    common = {"kernel": dom, "share_of_step": d["ms"] / tot_ms if tot_ms else None, "avg_launch_us": per_launch_ms * 1e3,
              "launches_per_step": d["launches"] / nprof,
              "timing": "CUDA events on the launching stream around every launch of a second pass over the same steps"}
    gemm_flops = sum(v["flops"] for k, v in prof.items() if k.startswith("gemm")) / nprof
    step_frac = {"algorithmic_TFLOP_per_step": gemm_flops / 1e12, "TFLOPs": gemm_flops / step_s / 1e12,
                 "frac_of_yardstick": gemm_flops / step_s / 1e12 / f64_peak, "frac_of_nominal_40": gemm_flops / step_s / 4e13,
                 "note": "all DMMA GEMM flops of a step (contraction + subspace-iteration panels; non-zero parity sectors "
                         "only) over the whole step time"}
    if dom.startswith("gemm") and d["flops"]:
        achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12
        roofline = dict({"bound": "tensor", "achieved": achieved, "peak": f64_peak, "unit": "TFLOP/s",
                         "frac": achieved / f64_peak, "traffic": None, "peak_source": f64_src,
                         "frac_of_nominal_40": achieved / 40.0,
                         "algorithmic_flops_per_launch": d["flops"] / max(d["launches"], 1)}, **common)
    else:
        achieved = (d["bytes"] / max(d["launches"], 1)) / (per_launch_ms * 1e-3) / 1e9 if d["bytes"] else 0.0
        roofline = dict({"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src}, **common)
    if (dom, chi) in NCU_TRAFFIC:
        roofline["traffic"], roofline["traffic_source"] = NCU_TRAFFIC[(dom, chi)]
    return roofline, shares, step_frac


# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the chi=32 step's kernel families from
# the committed `ncu --set full` captures: the working set is L2-resident, so DRAM traffic is far below the
# algorithmic bytes.  (For the HBM-bound kernel at scale, sign+permute D=128: 8.55 GB measured = 8.59 GB algorithmic,
# profiles/r1_sign_permute_D128_ncu_full.txt.)
NCU_TRAFFIC = {
    # the chi = 128 main contraction (8 sector GEMMs 4096 x 4096 x 8192 in one launch: 242 of the family's 247 ms per
    # step): 37.13 GB read + 2.35 GB written against 6.44 GB of operands + result -- with K = 8192 the A / B panels of
    # the 148 resident 128x64 tiles (2 x 1100 rows x 8192 x 16 B = 288 MB per wave, 111 waves) exceed L2, so every wave
    # streams its panels once; DRAM is 2 % busy, the tensor pipe 97.9 % (profiles/r2_tma_contraction_chi128_ncu_full.txt)
    ("gemm_tma_128x64", 128): (39483494000, "profiles/r2_tma_contraction_chi128_ncu_full.txt (dram read + write of the "
                                            "main-contraction launch, 98 % of the family's time)"),
    ("jacobi_persistent", 32): (1449984, "profiles/r1b_jacobi_persistent_ncu_full.txt"),
    ("gemm_skinny_32x32", 32): (304640, "profiles/r1_skinny_gemm_ncu_full.txt (32x32 panel configuration)"),
    ("chol_whiten", 32): (443136, "profiles/r1b_chol_whiten_ncu_full.txt"),
    ("gram_rotate", 32): (454400, "profiles/r1b_gram_rotate_ncu_full.txt"),
}


def fp64_tensor_peak(torch, dev, n=4096):
    """FP64 tensor-core yard-stick: cuBLAS ZGEMM n^3 timed here (library call used as the PEAK only)."""
    try:
        a = torch.randn(n, n, dtype=torch.complex128, device=dev)
        b = torch.randn(n, n, dtype=torch.complex128, device=dev)
        for _ in range(2):
            torch.matmul(a, b)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            torch.matmul(a, b)
        e.record()
        torch.cuda.synchronize()
        return 3 * 8.0 * n ** 3 / (s.elapsed_time(e) * 1e-3) / 1e12, "cuBLAS ZGEMM %d^3 measured in this run" % n
    except Exception:
        return 40.0, "nominal B200 FP64 tensor peak (cuBLAS yard-stick unavailable)"


def other_workloads(gtn, torch, data, stats, chi):
    """steps/s of the other coarse-graining drivers on the same Z2 tensor (not part of `value`):
    ATRG (example.py's default) in block and dense format, TRG in dense format, and TRG on a random
    Grassmann-even tensor (flat spectrum: the truncated SVD is rejected, full Jacobi SVD runs)."""
    import gtn_oracle as O
    g = gtn.gauge2d
    out = {}
    args = argparse.Namespace(chi=chi)

    def timed(fn, T, n=5, warm=2):
        """ms per step of the chain X -> fn(X) after `warm` steps (the chain goes on: no restart, the engine's
        iteration memory follows the drifting spectrum), adaptation frozen while timing"""
        X = T
        for _ in range(warm):
            X = fn(X)
        g.freeze(True)
        # settle: a step graph dropped just before the freeze is recorded again within a few steps; do not time that
        box = [X]

        def one():
            box[0] = fn(box[0])
        settle(g, one, max_steps=8)
        X = box[0]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            X = fn(X)
        torch.cuda.synchronize()
        g.freeze(False)
        return (time.perf_counter() - t0) / n * 1e3

    Td = g.zcap(gtn.dense(data, statistics=stats))
    Tb = Td.toblock()
    sat_b, sat_d = Tb, Td
    for _ in range(2):
        sat_b, _ = g.trg(sat_b, args.chi)
        sat_d, _ = g.trg(sat_d, args.chi)
    flip = [0]

    def atrg(X):
        flip[0] ^= 1
        return (g.atrg2dx if flip[0] else g.atrg2dy)(X, X, args.chi)[0]
    # long warm-up: the chain first has to drift from the TRG fixed point to the ATRG one (full-SVD fallbacks on
    # the way), then the iteration hints of the six decomposition sites settle and the step graphs are recorded
    out["atrg_block_chi%d_ms" % args.chi] = timed(atrg, sat_b, n=6, warm=31)
    out["atrg_dense_chi%d_ms" % args.chi] = timed(atrg, sat_d, n=6, warm=31)
    out["speculation"] = dict(g.SPEC_STATS)
    # the round-1 headline (chi = 32, same tensor every step, whole-step graph), device-resident
    for _ in range(24):
        g.trg(sat_b, args.chi)
    settle(g, lambda: g.trg(sat_b, args.chi))
    g.freeze(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        g.trg(sat_b, args.chi)
    torch.cuda.synchronize()
    out["trg_block_same_tensor_chi%d_ms" % args.chi] = (time.perf_counter() - t0) / 20 * 1e3
    g.freeze(False)
    out["trg_block_late_chain_chi%d_ms" % args.chi] = timed(lambda X: g.trg(X, args.chi)[0], sat_b, n=6, warm=30)
    out["trg_dense_chi%d_ms" % args.chi] = timed(lambda X: g.trg(X, args.chi)[0], sat_d, n=5, warm=30)
    rng = np.random.RandomState(3)
    R = O.random_dense((16, 16, 16, 16), (1, 1, -1, -1), dtype=complex, rng=rng)
    Rb = gtn.dense(R.data, statistics=R.statistics).toblock()
    out["trg_block_random_D16_chi16_ms"] = timed(lambda X: g.trg(X, 16)[0], Rb, n=2, warm=1)
    # flavour-direction HOTRG step (example.py --Nf 2): 6-leg tensors with bosonic legs, eig, hconjugate
    T6 = gtn.dense(data, statistics=stats)
    out["hotrg3dz_dense_Zcut%d_ms" % args.chi] = timed(lambda X: g.hotrg3dz(T6, T6, args.chi)[0], T6, n=2, warm=1)
    # BASELINE.json configs[2] "HOTRG chi=64": the same step at Zcut = 64 (parity with the reference to 1e-10 in
    # tests/test_z2_golden.py::test_gpu_hotrg3dz_z2_chi64_vs_reference; the reference takes 372 s for it on one core)
    out["hotrg3dz_z2_Zcut64_ms"] = timed(lambda X: g.hotrg3dz(T6, T6, 64)[0], T6, n=2, warm=1)
    out["einsum_sweep"] = einsum_sweep(gtn, torch, O)
    return out


def _written(r):
    """a timed einsum must have written its result: a pure permutation is otherwise returned unwritten
    (grassmanntn_b200._ops.LazyPermute) and would be timed as zero work"""
    bt = getattr(r, "_bt", None)
    if bt is not None:
        bt.buf
    elif hasattr(r, "buf"):
        r.buf
    return r


def einsum_sweep(gtn, torch, O):
    """BASELINE.json configs[4]: synthetic Grassmann einsums, 4-8 leg complex128 tensors (seeded,
    uniform [0,1) + i uniform [0,1), Grassmann-odd entries zeroed), sign-only and sign+contract.
    GB/s = 2*N*16 B algorithmic; TFLOP/s = 8*m*n*k over the non-zero parity sectors.  The same
    strings at 1/4 of the linear size are timed with the CPU oracle port for reference."""
    from grassmanntn_b200 import _engine as E
    cases = [
        ("ijkl->jkli", [((64,) * 4, (1, 1, -1, -1))]),
        ("ijkl->lkji", [((64,) * 4, (1, -1, 1, -1))]),
        ("abcdef->fedcba", [((16,) * 6, (1, 1, 1, -1, -1, -1))]),
        ("abcdefgh->hgfedcba", [((8,) * 8, (1, 1, 1, 1, -1, -1, -1, -1))]),
        ("ijkl,klmn->ijmn", [((64,) * 4, (1, 1, 1, 1)), ((64,) * 4, (-1, -1, 1, 1))]),
        ("abcdef,defghi->abcghi", [((16,) * 6, (1, 1, 1, 1, 1, 1)), ((16,) * 6, (-1, -1, -1, 1, 1, 1))]),
        ("abcd,cdef,efgh->abgh", [((32,) * 4, (1, 1, 1, 1)), ((32,) * 4, (-1, -1, 1, 1)), ((32,) * 4, (-1, -1, 1, 1))]),
        ("ijkl,klij", [((64,) * 4, (1, 1, 1, 1)), ((64,) * 4, (-1, -1, -1, -1))]),
        # the upper end of BASELINE.json's "dims 16-512": 1 GiB tensors with 512-wide fermionic legs
        ("ijkl->klij", [((512, 16, 512, 16), (1, 1, -1, -1))]),
        ("ijkl,klmn->ijmn", [((512, 16, 16, 512), (1, 1, 1, 1)), ((16, 512, 512, 16), (-1, -1, 1, 1))]),
    ]
    res = {}
    for sub, ops in cases:
        rng = np.random.RandomState(11)
        objs = []
        for shape, st in ops:
            if int(np.prod(shape)) > 1 << 24:
                # large cases: uniform [0,1) + i [0,1) drawn on the device (the CPU generator would take a minute),
                # Grassmann-odd entries removed by the block conversion's evenness trim
                x = torch.rand(shape, dtype=torch.float64, device="cuda").to(torch.complex128)
                x += 1j * torch.rand(shape, dtype=torch.float64, device="cuda")
                objs.append(gtn.trim_grassmann_odd(gtn.dense(x, statistics=st).toblock()))
                del x
                continue
            d = O.random_dense(shape, st, dtype=complex, rng=rng)
            objs.append(gtn.dense(d.data, statistics=st).toblock())
        for _ in range(2):
            r = _written(gtn.einsum(sub, *objs))
        torch.cuda.synchronize()
        E.PROF.start()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            r = _written(gtn.einsum(sub, *objs))
        e.record()
        torch.cuda.synchronize()
        pr = E.PROF.stop()
        ms = s.elapsed_time(e) / 5
        ent = {"ms": ms}
        if "sign_permute" in pr:
            ent["permute_GBps"] = pr["sign_permute"]["bytes"] / (pr["sign_permute"]["ms"] * 1e-3) / 1e9
        gm = [v for k, v in pr.items() if k.startswith("gemm") and v["flops"]]
        if gm:
            ent["gemm_TFLOPs"] = sum(v["flops"] for v in gm) / (sum(v["ms"] for v in gm) * 1e-3) / 1e12
            ent["gemm_family"] = [k for k in pr if k.startswith("gemm")]
        res[sub if sub not in res else "%s %s" % (sub, "x".join(str(x) for x in ops[0][0]))] = ent
    return res


def sharded_contraction(gtn, torch, dist, dev, D):
    """strong scaling of the TRG main contraction 'lxzk,jzxi->ijkl' at D = chi: output row tiles of the
    8 parity-block GEMMs sharded over the ranks, blocks completed with in-place NCCL all-gathers
    (grassmanntn_b200/parallel.py).  Device time, max over ranks."""
    from grassmanntn_b200 import parallel
    half = D // 2

    def even_block(stats, seed):
        gen = torch.Generator(device="cpu")
        gen.manual_seed(seed)
        b = gtn.zero_block_eo((half,) * 4, (half,) * 4, stats, dtype=complex)
        bt = b._bt
        for p_ in bt.patterns():
            if sum(p_) % 2 == 0:
                v = bt.block_view(p_)
                v.copy_(torch.view_as_complex(torch.rand(tuple(v.shape) + (2,), generator=gen, dtype=torch.float64)).to(dev))
            else:
                bt.zero.add(p_)
        return b
    VV, UU = even_block((1, 1, -1, 1), 1), even_block((-1, 1, -1, 1), 2)

    def timeit(n):
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            gtn.einsum('lxzk,jzxi->ijkl', VV, UU)
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    parallel.disable()
    timeit(1)
    ms1 = timeit(3)
    ref = gtn.einsum('lxzk,jzxi->ijkl', VV, UU)._bt.buf.clone()
    parallel.enable(min_flops=0.0, fused=False)
    timeit(1)
    msN = timeit(3)
    same_nccl = bool(torch.equal(gtn.einsum('lxzk,jzxi->ijkl', VV, UU)._bt.buf, ref))
    parallel.enable(min_flops=0.0, fused=True)
    timeit(1)
    fused_on = parallel.fused()
    msF = timeit(3)
    same_fused = bool(torch.equal(gtn.einsum('lxzk,jzxi->ijkl', VV, UU)._bt.buf, ref))
    parallel.disable()
    fl = 2.0 * D ** 6
    best = min(msN, msF) if fused_on else msN
    return {"D": D, "ranks": dist.get_world_size(), "ms_one_gpu": ms1, "ms_sharded": best, "speedup": ms1 / best,
            "TFLOPs_sharded_aggregate": fl / (best * 1e-3) / 1e12, "scaling": "strong",
            "ms_gemm_then_nccl_allgather": msN, "ms_fused_gemm_peer_store": msF if fused_on else None,
            "fused_available": fused_on, "bit_identical_to_one_gpu": {"nccl": same_nccl, "fused": same_fused}}


def microbench(gtn, E, torch, dev, args, hbm_peak):
    """large-size numbers for the two roofline kernels (not part of `value`)."""
    out = {}
    import gtn_oracle as O  # only for the seeded input generator
    # ---- sign + permute: dense 'ijkl->jkli' at D = 64 and 128 (complex128: 256 MiB / 4 GiB)
    for D in (64, 128):
        n = D ** 4
        x = torch.rand(n, dtype=torch.float64, device=dev).to(torch.complex128).view(D, D, D, D)
        A = gtn.dense(x, statistics=(1, 1, -1, -1))
        bt = A._get_bt()
        from grassmanntn_b200 import _ops
        for _ in range(3):
            r = _written(_ops.einsum_bt('ijkl->jkli', [bt]))
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10 if D == 64 else 3
        s.record()
        for _ in range(reps):
            r = _written(_ops.einsum_bt('ijkl->jkli', [bt]))
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
        gbs = 2 * n * 16 / (ms * 1e-3) / 1e9
        out["sign_permute_D%d" % D] = {"ms": ms, "GBps": gbs, "frac_hbm": gbs / hbm_peak,
                                       "bytes": 2 * n * 16, "note": "all 16 parity blocks, full (non-even) tensor"}
        del x, A, bt, r
    # ---- the TRG main contraction 'lxzk,jzxi->ijkl' (gauge2d.py:1734) on Grassmann-even block tensors:
    #      2 pack launches + 8 sector GEMMs in one grouped launch; algorithmic flops = 8*m*n*k over the
    #      non-zero parity sectors = 2*chi^6 (a quarter of the dense count)
    from grassmanntn_b200 import _ops as _o
    for D in (64, 128):
        try:
            np.random.seed(1)
            half = D // 2
            def even_block(stats):
                b = gtn.zero_block_eo((half,) * 4, (half,) * 4, stats, dtype=complex)
                bt = b._bt
                for p_ in bt.patterns():
                    if sum(p_) % 2 == 0:
                        v = bt.block_view(p_)
                        v.copy_(torch.rand(v.shape, dtype=torch.float64, device=dev) + 1j * torch.rand(v.shape, dtype=torch.float64, device=dev))
                    else:
                        bt.zero.add(p_)
                return b
            VV = even_block((1, 1, -1, 1))
            UU = even_block((-1, 1, -1, 1))
            for _ in range(2):
                r = _written(gtn.einsum('lxzk,jzxi->ijkl', VV, UU))
            torch.cuda.synchronize()
            E.PROF.start()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            s.record()
            for _ in range(reps):
                r = _written(gtn.einsum('lxzk,jzxi->ijkl', VV, UU))
            e.record()
            torch.cuda.synchronize()
            pr = E.PROF.stop()
            ms = s.elapsed_time(e) / reps
            flops = 2.0 * D ** 6
            fam = [k for k in pr if k.startswith("gemm")]
            gm = {"ms": sum(pr[k]["ms"] for k in fam), "flops": sum(pr[k]["flops"] for k in fam)}
            out["trg_contraction_D%d" % D] = {
                "ms_total": ms, "TFLOPs_total": flops / (ms * 1e-3) / 1e12,
                "ms_gemm": gm["ms"] / reps, "TFLOPs_gemm": gm["flops"] / reps / (gm["ms"] / reps * 1e-3) / 1e12,
                "ms_pack": pr["sign_permute"]["ms"] / reps,
                "pack_GBps": pr["sign_permute"]["bytes"] / reps / (pr["sign_permute"]["ms"] / reps * 1e-3) / 1e9,
                "algorithmic_flops": flops, "gemm_family": fam}
            del VV, UU, r
            torch.cuda.empty_cache()
        except Exception as ex:       # e.g. out of memory on a smaller part
            out["trg_contraction_D%d" % D] = {"error": repr(ex)[:200]}
    # ---- DMMA grouped GEMM vs cuBLAS ZGEMM yard-stick
    for N in (2048, 4096):
        a = torch.randn(N, N, dtype=torch.complex128, device=dev)
        b = torch.randn(N, N, dtype=torch.complex128, device=dev)
        for _ in range(2):
            c = E.gemm(a.view(-1), b.view(-1), N, N, N)
            c2 = a @ b
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            c = E.gemm(a.view(-1), b.view(-1), N, N, N)
        e.record()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for _ in range(3):
            c2 = a @ b
        e2.record()
        torch.cuda.synchronize()
        fl = 8.0 * N ** 3
        ours, cub = fl / (s.elapsed_time(e) / 3 * 1e-3) / 1e12, fl / (s2.elapsed_time(e2) / 3 * 1e-3) / 1e12
        err = float((c.view(N, N) - c2).abs().max() / c2.abs().max())
        out["zgemm_%d" % N] = {"ours_TFLOPs": ours, "cublas_zgemm_TFLOPs": cub, "frac_of_cublas": ours / cub,
                               "max_rel_diff_vs_cublas": err}
    return out


if __name__ == "__main__":
    main()
