"""gtn_sector_svd_trunc / gtn_sector_eigh_trunc / gtn_workspace_bytes (include/gtn_b200.h): the one-call C-ABI form of
the truncated sector decomposition (the driver loop of the subspace iteration, its certificate and the rank certificate
run inside the library) against numpy's LAPACK SVD on seeded matrices: singular values to 1e-10 s_0 (north_star),
the kept triplets through the gauge-invariant projector, rank rule on an exactly rank-deficient matrix, a flat
spectrum reported as GTN_ERR_NOT_CONVERGED, and the chi = 64 TRG chain with EVERY truncated decomposition routed through
the one-call (oracle golden)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _decaying(m, n, cplx, seed, rate=0.8, rank=None):
    rng = np.random.RandomState(seed)
    r = min(m, n) if rank is None else rank
    A = rng.randn(m, r) + (1j * rng.randn(m, r) if cplx else 0)
    B = rng.randn(r, n) + (1j * rng.randn(r, n) if cplx else 0)
    Qa, _ = np.linalg.qr(A)
    Qb, _ = np.linalg.qr(B.conj().T)
    s = rate ** np.arange(r)
    return (Qa * s) @ Qb.conj().T


def _onecall(torch, mats, ks, eig=False, robust=False):
    from grassmanntn_b200 import _cabi, _engine as E
    lib = _cabi.lib
    dev = torch.device("cuda")
    cplx = np.iscomplexobj(mats[0])
    dt = torch.complex128 if cplx else torch.float64
    nb, code = len(mats), E.dtype_code(dt)
    src = [torch.from_numpy(np.ascontiguousarray(m)).to(dev) for m in mats]
    m_ = (C.c_int64 * nb)(*[m.shape[0] for m in mats])
    n_ = (C.c_int64 * nb)(*[m.shape[1] for m in mats])
    k_ = (C.c_int32 * nb)(*ks)
    op = _cabi.GTN_OP_SECTOR_EIGH_TRUNC if eig else _cabi.GTN_OP_SECTOR_SVD_TRUNC
    nbytes = int(lib.gtn_workspace_bytes(op, code, nb, m_, n_, k_))
    assert nbytes > 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    U = [torch.empty(m.shape[0], k, dtype=dt, device=dev) for m, k in zip(mats, ks)]
    Vh = [torch.empty(k, m.shape[1], dtype=dt, device=dev) for m, k in zip(mats, ks)]
    ptrs = lambda ts: (C.c_void_p * nb)(*[t.data_ptr() for t in ts])
    S = (C.c_double * sum(ks))()
    lam = (C.c_double * (2 * sum(ks)))()
    rank = (C.c_int32 * nb)()
    info = _cabi.SvdInfo()
    info.robust = 1 if robust else 0
    if eig:
        rc = lib.gtn_sector_eigh_trunc(ptrs(src), m_, n_, nb, code, k_, 1e-14, ptrs(U), S, ptrs(Vh), lam, rank,
                                       E._ptr(ws), nbytes, C.byref(info), None)
    else:
        rc = lib.gtn_sector_svd_trunc(ptrs(src), m_, n_, nb, code, k_, 1e-14, ptrs(U), S, ptrs(Vh), rank, E._ptr(ws),
                                      nbytes, C.byref(info), None)
    torch.cuda.synchronize()
    s = np.frombuffer(S, dtype=np.float64).copy()
    lm = np.frombuffer(lam, dtype=np.float64).copy().view(np.complex128)
    outs, o = [], 0
    for b, k in enumerate(ks):
        outs.append((U[b].cpu().numpy(), s[o: o + k], Vh[b].cpu().numpy(), int(rank[b]), lm[o: o + k]))
        o += k
    return rc, outs, info


@pytest.mark.parametrize("cplx", [True, False])
def test_onecall_svd_vs_lapack(gtn, cplx):
    import torch
    mats = [_decaying(640, 512, cplx, 1), _decaying(512, 768, cplx, 2, rate=0.7), _decaying(512, 512, cplx, 3, rank=10)]
    ks = [16, 24, 16]
    rc, outs, info = _onecall(torch, mats, ks)
    assert rc == 0, rc
    assert info.launches > 0 and info.checks >= 1
    for M, k, (U, s, Vh, rank, _) in zip(mats, ks, outs):
        ref = np.linalg.svd(M, compute_uv=False)
        nnz = int(np.sum(ref / (ref[0] + 1e-14) > 1e-14))
        assert rank == min(k, nnz), (rank, nnz)
        assert np.abs(s[:rank] - ref[:rank]).max() <= 1e-10 * ref[0]
        Uk, Vk = U[:, :rank], Vh[:rank]
        assert np.abs(Uk.conj().T @ Uk - np.eye(rank)).max() <= 1e-10
        assert np.abs(Vk @ Vk.conj().T - np.eye(rank)).max() <= 1e-10
        # the kept triplets reproduce the best rank-`rank` approximation (Eckart-Young): compare the matrices
        Ur, sr, Vr = np.linalg.svd(M, full_matrices=False)
        best = (Ur[:, :rank] * sr[:rank]) @ Vr[:rank]
        assert np.abs((Uk * s[:rank]) @ Vk - best).max() <= 1e-9 * ref[0]


def test_onecall_flat_spectrum_is_reported(gtn):
    import torch
    rng = np.random.RandomState(5)
    M = rng.randn(512, 512) + 1j * rng.randn(512, 512)          # Marchenko-Pastur bulk: no gap at any cut
    rc, outs, info = _onecall(torch, [M], [16])
    from grassmanntn_b200 import _cabi
    assert rc == _cabi.GTN_ERR_NOT_CONVERGED


def test_onecall_steep_spectrum_needs_the_robust_mode(gtn):
    """singular values falling by 0.6 per index: s_40 = 2e-9 s_0 lies below what a Gram matrix resolves (3e-7), the plain
    run must refuse (its whitening dropped directions), the shifted-Cholesky-QR run must deliver all 40 to 1e-10 s_0"""
    import torch
    from grassmanntn_b200 import _cabi
    M = _decaying(512, 512, True, 21, rate=0.6)
    rc, outs, info = _onecall(torch, [M], [40])
    assert rc == _cabi.GTN_ERR_NOT_CONVERGED
    rc, outs, info = _onecall(torch, [M], [40], robust=True)
    assert rc == 0
    U, s, Vh, rank, _ = outs[0]
    ref = np.linalg.svd(M, compute_uv=False)
    assert rank == 40 and np.abs(s - ref[:40]).max() <= 1e-10 * ref[0]
    assert np.abs(s[24:] / ref[24:40] - 1).max() <= 1e-4            # the tail (7e-6 ... 2e-9 s_0) also in relative terms


def test_onecall_eigh_vs_numpy(gtn):
    import torch
    rng = np.random.RandomState(7)
    n = 384
    Q, _ = np.linalg.qr(rng.randn(n, n) + 1j * rng.randn(n, n))
    lam = 0.75 ** np.arange(n) * np.where(np.arange(n) % 3 == 1, -1.0, 1.0)
    H = (Q * lam) @ Q.conj().T
    H = 0.5 * (H + H.conj().T)
    rc, outs, info = _onecall(torch, [H], [12], eig=True)
    assert rc == 0
    U, s, Vh, rank, lm = outs[0]
    assert rank == 12
    assert np.abs(np.sort(np.abs(lm))[::-1] - np.abs(lam[:12])).max() <= 1e-10
    assert np.abs(lm.imag).max() <= 1e-10
    assert np.array_equal(np.sign(lm.real), np.sign(lam[:12]))


def test_chi64_chain_through_the_onecall(gtn, monkeypatch):
    """every truncated decomposition of the chi = 64 chain (incl. the rank-deficient second step) behind ONE C call"""
    from grassmanntn_b200 import _engine as E
    monkeypatch.setattr(E, "TRUNC_PLAN_CACHE_BYTES", 0)
    ref = np.load(os.path.join(G, "chains.npz"))["oracle_trg_chi64"]
    g = gtn.gauge2d
    before = dict(E.ONE_CALL_STATS)
    T = g.zcap(g.load_initial_tensor().toblock())
    T, recs = g.coarse_grain(T, cgsteps=4, dcut=64, method="trg", boundary_conditions="anti-periodic")
    for i in range(1, 5):
        r = recs[i]
        assert tuple(r["shape"]) == (int(ref[i - 1, 3]), int(ref[i - 1, 4])), (i, r["shape"])
        assert abs(r["Tnorm"] - ref[i - 1, 0]) <= 1e-10 * ref[i - 1, 0], (i, r["Tnorm"], ref[i - 1, 0])
        assert abs(r["F"] - complex(ref[i - 1, 1], ref[i - 1, 2])) <= 1e-10 * abs(r["F"]), (i, r["F"])
    assert E.ONE_CALL_STATS["accepted"] >= before["accepted"] + 2
