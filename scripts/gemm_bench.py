"""GEMM tile configurations side by side on the shapes of the coarse-graining steps (run on the GPU box):
cp.async 64x64 (config 0), skinny 32x32 (1), TMA 64x64 (4), TMA 128x64 (12) and cuBLAS ZGEMM as the yard-stick.
Writes gpurun_out/gemm_bench.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grassmanntn_b200 as gtn  # noqa: F401
from grassmanntn_b200 import _engine as E

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts)


def case(name, m, n, k, ngroups, dt=torch.complex128, configs=(0, 1, 4, 12), cublas=True):
    A = torch.randn(ngroups * m * k, dtype=dt, device=dev)
    B = torch.randn(ngroups * k * n, dtype=dt, device=dev)
    Cm = torch.empty(ngroups * m * n, dtype=dt, device=dev)
    groups = [dict(a_off=i * m * k, b_off=i * k * n, c_off=i * m * n, lda=k, ldb=n, ldc=n, m=m, n=n, k=k) for i in range(ngroups)]
    fl = (8 if dt == torch.complex128 else 2) * m * n * k * ngroups
    out = {"m": m, "n": n, "k": k, "groups": ngroups, "dtype": str(dt), "GFLOP": fl / 1e9}
    ref = None
    for cfg in configs:
        try:
            plan = E.GemmPlan(groups, dt, config=cfg)
            ms = timeit(lambda: plan.run(A, B, Cm))
            out["cfg%d_ms" % cfg] = ms
            out["cfg%d_TFLOPs" % cfg] = fl / (ms * 1e-3) / 1e12
            if ref is None:
                ref = Cm.clone()
            else:
                out["cfg%d_maxdiff_vs_first" % cfg] = float((Cm - ref).abs().max())
        except Exception as ex:
            out["cfg%d_error" % cfg] = repr(ex)[:200]
    if cublas:
        a3, b3 = A.view(ngroups, m, k), B.view(ngroups, k, n)
        ms = timeit(lambda: torch.bmm(a3, b3))
        out["cublas_ms"] = ms
        out["cublas_TFLOPs"] = fl / (ms * 1e-3) / 1e12
        if ref is not None:
            out["maxdiff_vs_cublas"] = float((ref.view(ngroups, m, n) - torch.bmm(a3, b3)).abs().max())
    print(name, json.dumps(out), flush=True)
    return out


res = {}
res["zgemm_2048"] = case("zgemm_2048", 2048, 2048, 2048, 1)
res["zgemm_4096"] = case("zgemm_4096", 4096, 4096, 4096, 1)
res["dgemm_4096"] = case("dgemm_4096", 4096, 4096, 4096, 1, dt=torch.float64)
res["panel_chi128"] = case("panel_chi128 (l=128 x 8192 x 8192, 4 sectors)", 128, 8192, 8192, 4)
res["panel_chi64"] = case("panel_chi64 (l=64 x 2048 x 2048, 4 sectors)", 64, 2048, 2048, 4, configs=(0, 1, 4))
res["gram_chi128"] = case("gram_chi128 (128 x 128 x 2048, 16 slices)", 128, 128, 2048, 16)
res["apply_chi128"] = case("apply_chi128 (128 x 8192 x 128, 4)", 128, 8192, 128, 4)
res["contraction_chi64"] = case("contraction chi=64 (2048^3, 8 blocks)", 2048, 2048, 2048, 8, configs=(0, 12))
res["contraction_chi128_1blk"] = case("contraction chi=128 (8192^3, 1 of 8 blocks)", 8192, 8192, 8192, 1, configs=(0, 12), cublas=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)
