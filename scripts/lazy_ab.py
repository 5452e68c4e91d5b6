"""A/B of the unwritten-permutation path (_ops.LAZY_PERMUTE) in one process on one GPU: ms per step and launches per
step of TRG (chi = 32 step graph, 64, 128), ATRG chi = 32 and hotrg3dz Zcut 64 on the Z2 tensor, with every leg
permutation written (0) and with consumers packing from the stored source (1).  Run on the GPU box."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import gauge2d as g, _ops, _engine as E

sizes = [int(x) for x in os.environ.get("CHIS", "32,64,128").split(",")]
out = {}


def reset(lazy):
    _ops.LAZY_PERMUTE = bool(lazy)
    g._drop_step_graphs()
    _ops._einsum_cache.clear()
    E._plan_cache.clear()
    E.drop_graphs()
    for k in _ops.LAZY_STATS:
        _ops.LAZY_STATS[k] = 0


def timed(fn, n):
    torch.cuda.synchronize()
    c0 = gtn.launch_count()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return round((time.perf_counter() - t0) / n * 1e3, 3), (gtn.launch_count() - c0) / n


def saturate(T, chi):
    n = 0
    while tuple(T.effective_shape) != (chi,) * 4 and n < 8:
        T, _ = g.trg(T, chi)
        n += 1
    return T


T0 = g.zcap(g.load_initial_tensor()).toblock()
T6 = g.load_initial_tensor()
for lazy in (0, 1, 0, 1):
    tag = "lazy%d" % lazy
    rec = out.setdefault(tag, {})
    reset(lazy)
    for chi in sizes:
        T = saturate(T0, chi)
        for _ in range(24 if chi <= 32 else 4):
            g.trg(T, chi)
        g.freeze(True)
        for _ in range(3):
            g.trg(T, chi)
        ms, nl = timed(lambda: g.trg(T, chi), 20 if chi <= 32 else (6 if chi <= 64 else 3))
        g.freeze(False)
        rec.setdefault("trg_chi%d_ms" % chi, []).append(ms)
        rec["trg_chi%d_launches" % chi] = nl
        rec["trg_chi%d_Tnorm" % chi] = float(g.trg(T, chi)[1])
    flip = [0]
    box = [saturate(T0, 32)]

    def atrg():
        flip[0] ^= 1
        box[0] = (g.atrg2dx if flip[0] else g.atrg2dy)(box[0], box[0], 32)[0]
    for _ in range(34):
        atrg()
    g.freeze(True)
    for _ in range(4):
        atrg()
    ms, nl = timed(atrg, 6)
    g.freeze(False)
    rec.setdefault("atrg_chi32_ms", []).append(ms)
    rec["atrg_chi32_launches"] = nl
    rec["atrg_chi32_Tnorm_after_45_steps"] = float(g.atrg2dy(box[0], box[0], 32)[1])
    rec["step_graphs"] = {k: v for k, v in g.STEP_GRAPH_STATS.items() if isinstance(v, int)}
    for _ in range(2):
        g.hotrg3dz(T6, T6, 64)
    ms, nl = timed(lambda: g.hotrg3dz(T6, T6, 64), 5)
    rec.setdefault("hotrg3dz_zcut64_ms", []).append(ms)
    rec["hotrg3dz_zcut64_launches"] = nl
    rec["hotrg3dz_zcut64_Tnorm"] = float(g.hotrg3dz(T6, T6, 64)[1])
    rec["lazy_stats"] = dict(_ops.LAZY_STATS)
    rec["peak_mem_GiB"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
    torch.cuda.reset_peak_memory_stats()
    print(tag, json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r2i_lazy_ab.json", "w"), indent=1)
