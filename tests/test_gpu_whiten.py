"""gtn_chol_whiten (pivoted Cholesky whitening of the l x l Gram matrices of the subspace iteration) in its three
variants -- register tiles (n <= 80), packed lower triangle in shared memory (80 < n <= 128, chi = 128 runs), global
scratch (n > 128) -- against numpy: T G T^H = I on the detected rank, zero rows beyond it, rank detection on a
rank-deficient matrix, split-K partial sums, complex128 and float64.  Tolerance eps * cond(G) (numpy's own Cholesky
whitening of the same matrices lands within a factor 50 below that), floor 1e-9."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _whiten(torch, G_list, nsplit=1, rel_thr=1e-13):
    from grassmanntn_b200 import _engine as E
    from grassmanntn_b200._cabi import check, lib
    dev = torch.device("cuda")
    cplx = np.iscomplexobj(G_list[0])
    dt = torch.complex128 if cplx else torch.float64
    nb = len(G_list)
    ns = [g.shape[-1] for g in G_list]
    g_off, t_off, acc_g, acc_t = [], [], 0, 0
    for g, n in zip(G_list, ns):
        g_off.append(acc_g)
        t_off.append(acc_t)
        acc_g += nsplit * n * n
        acc_t += n * n
    Gbuf = torch.zeros(acc_g, dtype=dt, device=dev)
    for g, n, off in zip(G_list, ns, g_off):
        parts = np.zeros((nsplit, n, n), dtype=g.dtype)
        parts[0] = g
        if nsplit > 1:                      # the slices sum to G
            rng = np.random.RandomState(5)
            for sp in range(1, nsplit):
                d = rng.rand(n, n) + (1j * rng.rand(n, n) if cplx else 0)
                parts[sp] = d
                parts[0] = parts[0] - d
        Gbuf[off: off + nsplit * n * n] = torch.from_numpy(parts.reshape(-1)).to(dev)
    Tbuf = torch.full((acc_t,), 7.0, dtype=dt, device=dev)
    g_off_d = torch.tensor(g_off, dtype=torch.int64, device=dev)       # (kept alive until the kernel has run)
    t_off_d = torch.tensor(t_off, dtype=torch.int64, device=dev)
    n_d = torch.tensor(ns, dtype=torch.int32, device=dev)
    kept = torch.zeros(nb, dtype=torch.int32, device=dev)
    se = int(lib.gtn_chol_whiten_scratch_elems(max(ns)))
    scratch = torch.empty(max(se * nb, 1), dtype=torch.complex128, device=dev)
    check(lib.gtn_chol_whiten(E._ptr(Gbuf), E._ptr(Tbuf), E.dtype_code(dt), E._ptr(g_off_d), E._ptr(t_off_d),
                              E._ptr(n_d), nb, max(ns), nsplit, rel_thr, E._ptr(kept), E._ptr(scratch), None),
          "gtn_chol_whiten")
    torch.cuda.synchronize()
    Ts = [Tbuf[o: o + n * n].view(n, n).cpu().numpy() for o, n in zip(t_off, ns)]
    return Ts, kept.cpu().numpy()


def _gram(n, rank, cplx, seed, cond=1e6):
    rng = np.random.RandomState(seed)
    Y = rng.randn(rank, 4 * n) + (1j * rng.randn(rank, 4 * n) if cplx else 0)
    Y *= np.logspace(0, -np.log10(cond) / 2, rank)[:, None]
    M = rng.randn(n, rank) + (1j * rng.randn(n, rank) if cplx else 0)
    X = M @ Y                                            # n x 4n of rank `rank`
    return X @ X.conj().T


@pytest.mark.parametrize("cplx", [True, False])
@pytest.mark.parametrize("n", [24, 48, 80, 96, 128, 160, 256])
def test_chol_whiten_variants(gtn, n, cplx):
    import torch
    for rank, nsplit in ((n, 1), (n, 4), (max(n - 7, 1), 1)):
        G = _gram(n, rank, cplx, 100 + n + rank)
        G2 = _gram(n, rank, cplx, 300 + n + rank)       # two problems per launch
        Ts, kept = _whiten(torch, [G, G2], nsplit=nsplit)
        for T, Gm, r in zip(Ts, (G, G2), kept):
            assert r == rank, (n, rank, nsplit, r)
            W = T @ Gm @ T.conj().T
            lam = np.linalg.eigvalsh(Gm)[::-1]
            tol = max(1e-9, 2.3e-16 * lam[0] / lam[r - 1])     # a Cholesky whitening is good to eps * cond(G)
            err = np.abs(W[:r, :r] - np.eye(r)).max()
            assert err <= tol, (n, rank, nsplit, err, tol)
            if r < n:
                assert np.abs(T[r:, :]).max() == 0.0
