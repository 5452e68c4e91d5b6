"""TMA-staged DMMA GEMM (csrc/gtn_gemm_tma.cu, gtn_grouped_gemm_tma) against a plain torch float64 / complex128
matmul of the same operands and against the cp.async kernel (gtn_grouped_gemm): ragged extents (TMA zero-fill of
partial boxes, boxes entirely out of range), K tails, leading dimensions larger than the extents, several groups
with offsets in one launch, batch strides, beta = 1 accumulation, alpha = -1 (block sign), both tile configurations,
and the rasterised tile walk with a band that does not divide the tile grid.
Tolerance: 1e-13 relative to |A| |B| k (float64 accumulation in a different order than torch's)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(gtn, torch, groups, A, B, Cm, dtype, config):
    from grassmanntn_b200 import _engine as E
    plan = E.GemmPlan(groups, dtype, config=config)
    plan.run(A, B, Cm)
    torch.cuda.synchronize()
    return plan


def _ref(torch, g, A, B, C0, bi=0):
    a = torch.as_strided(A, (g["m"], g["k"]), (g["lda"], 1), g["a_off"] + bi * g.get("bsa", 0))
    b = torch.as_strided(B, (g["k"], g["n"]), (g["ldb"], 1), g["b_off"] + bi * g.get("bsb", 0))
    r = g.get("alpha", 1.0) * (a @ b)
    if g.get("beta", 0.0):
        r = r + g["beta"] * torch.as_strided(C0, (g["m"], g["n"]), (g["ldc"], 1), g["c_off"] + bi * g.get("bsc", 0))
    return r


SHAPES = [(128, 64, 8), (128, 64, 9), (64, 64, 64), (200, 130, 77), (257, 65, 130), (64, 200, 31), (129, 72, 256),
          (1100, 90, 40)]


@pytest.mark.parametrize("cplx", [True, False])
@pytest.mark.parametrize("config", [12, 4])
def test_tma_gemm_shapes(gtn, cplx, config):
    import torch
    dt = torch.complex128 if cplx else torch.float64
    gen = torch.Generator(device="cuda").manual_seed(7)
    for (m, n, k) in SHAPES:
        lda, ldb, ldc = k + (6 if cplx else 6), n + 4, n + 2           # even paddings: float64 rows stay 16-byte aligned
        if not cplx:
            lda += lda & 1
            ldb += ldb & 1
        A = torch.randn(m * lda + 64, dtype=dt, device="cuda", generator=gen)
        B = torch.randn(k * ldb + 64, dtype=dt, device="cuda", generator=gen)
        C0 = torch.randn(m * ldc + 64, dtype=dt, device="cuda", generator=gen)
        for alpha, beta in ((1.0, 0.0), (-1.0, 1.0)):
            g = dict(a_off=2, b_off=4, c_off=3, lda=lda, ldb=ldb, ldc=ldc, m=m, n=n, k=k, alpha=alpha, beta=beta)
            Cm = C0.clone()
            _run(gtn, torch, [g], A, B, Cm, dt, config)
            got = torch.as_strided(Cm, (m, n), (ldc, 1), 3)
            ref = _ref(torch, g, A, B, C0)
            tol = 1e-13 * k * float(A.abs().max() * B.abs().max())
            assert float((got - ref).abs().max()) <= tol, (m, n, k, alpha, beta, config, cplx)
            # nothing outside the m x n window was touched
            mask = torch.ones_like(Cm, dtype=torch.bool)
            torch.as_strided(mask, (m, n), (ldc, 1), 3).fill_(False)
            assert torch.equal(Cm[mask], C0[mask])
            # same numbers as the cp.async kernel (same K order inside a DMMA chain)
            C2 = C0.clone()
            _run(gtn, torch, [g], A, B, C2, dt, 0)
            assert float((Cm - C2).abs().max()) <= tol


@pytest.mark.parametrize("cplx", [True, False])
def test_tma_gemm_groups_batch_raster(gtn, cplx):
    import torch
    dt = torch.complex128 if cplx else torch.float64
    gen = torch.Generator(device="cuda").manual_seed(11)
    # three groups of different shapes + one batched group, one launch; 19 row tiles of 64 (band of 16 + a band of 3)
    specs = [(1200, 96, 72, 1), (64, 64, 200, 1), (130, 260, 64, 1), (96, 80, 48, 3)]
    groups, ao, bo, co = [], 0, 0, 0
    for m, n, k, batch in specs:
        groups.append(dict(a_off=ao, b_off=bo, c_off=co, lda=k, ldb=n, ldc=n, m=m, n=n, k=k, batch=batch,
                           bsa=m * k, bsb=k * n, bsc=m * n))
        ao += m * k * batch
        bo += k * n * batch
        co += m * n * batch
    A = torch.randn(ao, dtype=dt, device="cuda", generator=gen)
    B = torch.randn(bo, dtype=dt, device="cuda", generator=gen)
    for config in (12, 4):
        Cm = torch.zeros(co, dtype=dt, device="cuda")
        plan = _run(gtn, torch, groups, A, B, Cm, dt, config)
        assert plan.config == config and plan.family.startswith("gemm_tma")
        for g in groups:
            for bi in range(g["batch"]):
                got = torch.as_strided(Cm, (g["m"], g["n"]), (g["ldc"], 1), g["c_off"] + bi * g["bsc"])
                ref = _ref(torch, g, A, B, Cm, bi)
                assert float((got - ref).abs().max()) <= 1e-13 * g["k"] * float(A.abs().max() * B.abs().max())


def test_tma_plan_selection(gtn):
    """large aligned products take the TMA tiles; products whose tile grid leaves a partial last wave, ragged /
    misaligned / skinny ones take the fine 32x32 or the 64x64 cp.async tiles (GemmPlan._choose)"""
    import torch
    from grassmanntn_b200 import _engine as E
    big = dict(a_off=0, b_off=0, c_off=0, lda=4096, ldb=4096, ldc=4096, m=4096, n=4096, k=4096)
    assert E.GemmPlan([big], torch.complex128).config == 12
    assert E.GemmPlan([dict(big, m=2048, n=2048, k=2048)], torch.complex128).config == 1     # 3.5 waves of 128x64 tiles
    odd = dict(big, lda=4097)
    assert E.GemmPlan([odd], torch.float64).config in (0, 1)       # float64 rows not 16-byte aligned: no TMA
    assert E.GemmPlan([odd], torch.complex128).config == 12
    small = dict(a_off=0, b_off=0, c_off=0, lda=13, ldb=12, ldc=12, m=13, n=12, k=13)
    assert E.GemmPlan([small], torch.complex128).config == 1
