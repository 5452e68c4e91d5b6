"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Tolerances: element-wise data produced by sign/permute work must be BIT-EXACT (0.0);
contractions 1e-12 relative (different summation order); singular values / Tnorm / free energy
1e-10 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

import gtn_oracle as O

pytestmark = pytest.mark.gpu

EINSUM_CASES = [
    ('ijkl->jkli', [((4, 4, 4, 4), (1, 1, -1, -1))]),
    ('ijkl->klij', [((4, 2, 8, 4), (1, -1, -1, 1))]),
    ('ijkl->lijk', [((8, 4, 8, 4), (1, 1, -1, -1))]),
    ('ijkl->jikl', [((8, 4, 8, 4), (1, 1, -1, -1))]),
    ('lai->ila', [((8, 4, 8), (1, -1, 1))]),
    ('ijkl,klmn->ijmn', [((4, 4, 4, 2), (1, 1, 1, 1)), ((4, 2, 4, 8), (-1, -1, 1, -1))]),
    ('kwz,lxw->lxzk', [((4, 4, 8), (1, -1, 1)), ((4, 2, 4), (-1, 1, 1))]),
    ('yxi,zyj->jzxi', [((4, 2, 8), (1, -1, 1)), ((4, 4, 2), (1, -1, -1))]),
    ('lxzk,jzxi->ijkl', [((4, 2, 8, 4), (1, 1, -1, 1)), ((2, 8, 2, 4), (-1, 1, -1, 1))]),
    ('ijij', [((4, 2, 4, 2), (1, 1, -1, -1))]),
    ('ijij', [((4, 2, 4, 2), (-1, 1, 1, -1))]),
    ('ijkl,klij', [((4, 2, 4, 8), (1, 1, -1, -1)), ((4, 8, 4, 2), (1, 1, -1, -1))]),
    ('IJIK,iKiJ', [((4, 2, 4, 8), (1, 1, -1, -1)), ((2, 8, 2, 2), (1, 1, -1, -1))]),
    ('IJKLij,ij->IJKL', [((4, 2, 4, 2, 3, 3), (1, 1, -1, -1, 0, 0)), ((3, 3), (0, 0))]),
    ('i1 i3 a, j1 j3 b -> i1 i3 ab j1 j3', [((4, 2, 3), (1, -1, 0)), ((2, 4, 2), (1, -1, 0))]),
    ('t s al m , lbm -> t s ab m', [((4, 2, 4, 2, 3), (1, -1, 1, 1, 0)), ((2, 4, 3), (-1, 1, 0))]),
    ('abx,xc->abc', [((4, 4, 8), (1, 1, 1)), ((8, 8), (-1, 1))]),
    ('ax,xbc->abc', [((8, 8), (-1, 1)), ((8, 4, 4), (-1, -1, -1))]),
    ('ajk,jib->aibk', [((4, 2, 8), (-1, -1, -1)), ((2, 4, 4), (1, 1, 1))]),
    ('iax,xbj->ijab', [((4, 2, 8), (1, 1, 1)), ((8, 4, 2), (-1, 1, -1))]),
    ('ab,bc,cd->ad', [((4, 8), (1, 1)), ((8, 2), (-1, 1)), ((2, 4), (-1, -1))]),
    ('abc,dbe,fce->adf', [((4, 8, 2), (1, 1, 1)), ((2, 8, 4), (1, -1, 1)), ((4, 2, 4), (-1, -1, -1))]),
    ('IJIJmn,KLKLmn->mn', [((4, 2, 4, 2, 3, 2), (1, 1, -1, -1, 0, 0)), ((2, 2, 2, 2, 3, 2), (1, 1, -1, -1, 0, 0))]),
    ('xa,xb->ab', [((4, 2), (1, 1)), ((4, 8), (-1, -1))]),
]


def _mk(gtn, shape, stats, rng, cplx=True, trim=False):
    o = O.random_dense(shape, stats, dtype=complex if cplx else float, rng=rng, skip_trimming=not trim)
    return o, gtn.dense(o.data, statistics=stats)


def _relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("trim", [False, True])
@pytest.mark.parametrize("case", range(len(EINSUM_CASES)))
def test_einsum_vs_oracle(gtn, case, trim):
    sub, ops = EINSUM_CASES[case]
    rng = np.random.RandomState(1000 + case)
    pairs = [_mk(gtn, s, st, rng, trim=trim) for s, st in ops]
    ref = O.einsum(sub, *[p[0] for p in pairs])
    got = gtn.einsum(sub, *[p[1] for p in pairs])
    if isinstance(ref, O.Dense):
        assert tuple(got.statistics) == tuple(ref.statistics)
        assert got.shape == ref.shape
        err = _relerr(got.data.cpu().numpy(), ref.data)
        single = len(ops) == 1 and all(sub.split('->')[0].count(c) == 1 for c in sub.split('->')[0])
        assert err == 0.0 if single else err < 1e-12, (sub, err)
    else:
        assert abs(got - ref) <= 1e-12 * max(abs(ref), 1.0), (sub, got, ref)


def test_einsum_real_dtype(gtn):
    rng = np.random.RandomState(5)
    a, A = _mk(gtn, (4, 8, 4), (1, 1, -1), rng, cplx=False)
    b, B = _mk(gtn, (4, 8, 2), (1, -1, -1), rng, cplx=False)
    ref = O.einsum('abc,cbd->ad', a, b)
    got = gtn.einsum('abc,cbd->ad', A, B)
    assert got.data.dtype.is_floating_point
    assert _relerr(got.data.cpu().numpy(), ref.data) < 1e-12


def test_format_encoder_switches_bit_exact(gtn):
    rng = np.random.RandomState(7)
    a, A = _mk(gtn, (4, 4, 2, 8), (1, -1, -1, 1), rng)
    for fn in ("switch_format", "switch_encoder", "switch_parity"):
        ref = getattr(a, fn)()
        got = getattr(A, fn)()
        assert (got.format, got.encoder) == (ref.format, ref.encoder)
        assert np.array_equal(got.data.cpu().numpy(), ref.data), fn
    ref = a.switch_format().switch_encoder()
    got = A.switch_format().switch_encoder()
    assert np.array_equal(got.data.cpu().numpy(), ref.data)
    r2 = O.einsum('ijkl->lkji', ref)
    g2 = gtn.einsum('ijkl->lkji', got)
    assert (g2.format, g2.encoder) == (r2.format, r2.encoder)
    assert np.array_equal(g2.data.cpu().numpy(), r2.data)


@pytest.mark.parametrize("cut", [None, 8, 6])
def test_svd_vs_oracle(gtn, cut):
    rng = np.random.RandomState(11)
    a, A = _mk(gtn, (4, 4, 4, 4), (1, 1, -1, -1), rng, trim=True)
    Ur, Sr, Vr = O.svd(a, 'ab|cd', cut)
    U, S, V = A.svd('ab|cd', cut)
    assert S.shape == Sr.shape and U.shape == Ur.shape and V.shape == Vr.shape
    assert tuple(U.statistics) == tuple(Ur.statistics) and tuple(V.statistics) == tuple(Vr.statistics)
    sr = np.sort(np.abs(np.diag(Sr.data)))[::-1]
    sg = np.sort(np.abs(np.diag(S.data.cpu().numpy())))[::-1]
    assert np.abs(sr - sg).max() <= 1e-10 * sr[0]
    rec_r = O.einsum('abx,xy,ycd->abcd', Ur, Sr, Vr).data
    rec_g = gtn.einsum('abx,xy,ycd->abcd', U, S, V).data.cpu().numpy()
    assert _relerr(rec_g, rec_r) < 1e-10
    if cut is None:
        assert _relerr(rec_g, a.data) < 1e-12


def test_svd_with_bosonic_legs(gtn):
    rng = np.random.RandomState(12)
    a, A = _mk(gtn, (4, 2, 3, 4, 2), (1, -1, 0, -1, 1), rng, trim=True)
    Ur, Sr, Vr = O.svd(a, 'abc|de', None)
    U, S, V = A.svd('abc|de', None)
    rec_g = gtn.einsum('abcx,xy,yde->abcde', U, S, V).data.cpu().numpy()
    assert _relerr(rec_g, a.data) < 1e-12
    assert tuple(U.statistics) == tuple(Ur.statistics)


def test_hconjugate_bit_exact(gtn):
    rng = np.random.RandomState(13)
    for shape, stats, s in [((4, 2, 3, 4, 2), (1, -1, 0, -1, 1), 'abc|de'), ((4, 4, 4, 4), (1, 1, -1, -1), 'ab|cd'),
                            ((4, 8, 2), (-1, 1, 1), 'a|bc')]:
        a, A = _mk(gtn, shape, stats, rng)
        ref = O.hconjugate(a, s)
        got = A.hconjugate(s)
        assert tuple(got.statistics) == tuple(ref.statistics)
        assert np.array_equal(got.data.cpu().numpy(), ref.data), s


def test_eig_vs_oracle(gtn):
    rng = np.random.RandomState(14)
    a, A = _mk(gtn, (4, 4, 4, 4), (1, 1, -1, -1), rng, trim=True)
    Mr = O.einsum('abcd,cdef->abef', O.hconjugate(a, 'ab|cd'), a)
    Mg = gtn.einsum('abcd,cdef->abef', A.hconjugate('ab|cd'), A)
    assert _relerr(Mg.data.cpu().numpy(), Mr.data) < 1e-12
    Ur, Sr, Vr = O.eig(Mr, 'ab|cd', 8)
    U, S, V = Mg.eig('ab|cd', 8)
    sr = np.sort(np.abs(np.diag(Sr.data)))[::-1]
    sg = np.sort(np.abs(np.diag(S.data.cpu().numpy())))[::-1]
    assert np.abs(sr - sg).max() <= 1e-10 * sr[0]
    rec_r = O.einsum('abx,xy,ycd->abcd', Ur, Sr, Vr).data
    rec_g = gtn.einsum('abx,xy,ycd->abcd', U, S, V).data.cpu().numpy()
    assert _relerr(rec_g, rec_r) < 1e-10


@pytest.mark.parametrize("fmt", ["dense", "block"])
@pytest.mark.parametrize("algo", ["trg", "atrg2dy", "atrg2dx"])
def test_cg_step_vs_oracle(gtn, algo, fmt):
    rng = np.random.RandomState(21)
    a, A = _mk(gtn, (4, 4, 4, 4), (1, 1, -1, -1), rng, trim=True)
    cut = 8 if fmt == "dense" else 6
    rule = fmt
    if fmt == "block":
        A = A.toblock()
    g = gtn.gauge2d
    if algo == "trg":
        Tr, nr, er = O.trg(a, cut, rule=rule, error_test=True)
        Tg, ng, eg = g.trg(A, cut, error_test=True)
    else:
        fo = getattr(O, algo)
        fg = getattr(g, algo)
        Tr, nr, er = fo(a, a, cut, rule=rule, error_test=True)
        Tg, ng, eg = fg(A, A, cut, error_test=True)
    assert abs(ng - nr) <= 1e-10 * nr, (ng, nr)
    assert abs(eg - er) <= 1e-9 * max(er, 1e-3)
    Fr = O.logZ(Tr, 'anti-periodic', block_format=(fmt == "block"))
    Fg = g.logZ(Tg, 'anti-periodic')
    assert abs(Fg - Fr) <= 1e-10 * max(abs(Fr), 1.0)


def _decaying_even_tensor(D, rng, rate=0.5):
    """Grassmann-even (D,D,D,D) tensor whose (ab|cd) sector matrices have a geometric spectrum."""
    a = O.random_dense((D, D, D, D), (1, 1, -1, -1), dtype=complex, rng=rng)
    U, S, V = O.svd(a, 'ab|cd')
    d = S.data.copy()
    n = d.shape[0]
    for i in range(n):
        d[i, i] = d[i, i] / max(abs(d[i, i]), 1e-300) * np.exp(-rate * (i // 2))
    S2 = O.Dense(d, S.statistics, S.encoder, S.format)
    return O.einsum('abx,xy,ycd->abcd', U, S2, V)


@pytest.mark.parametrize("kind", ["decaying", "flat"])
def test_truncated_svd_path(gtn, kind):
    """the randomized subspace path (accepted for decaying spectra, rejected -> full Jacobi for flat
    ones) must give the same kept singular values as the oracle's full LAPACK SVD"""
    from grassmanntn_b200 import _ops
    rng = np.random.RandomState(31)
    D = 16
    a = _decaying_even_tensor(D, rng) if kind == "decaying" else O.random_dense((D,) * 4, (1, 1, -1, -1), dtype=complex, rng=rng)
    A = gtn.dense(a.data, statistics=(1, 1, -1, -1))
    before = dict(_ops.SVD_PATH_STATS)
    for cut, fmt in ((8, "dense"), (10, "block")):
        Ur, Sr, Vr, counts = O.svd(a, 'ab|cd', cut, rule=fmt, return_counts=True)
        X = A if fmt == "dense" else A.toblock()
        U, S, V = X.svd('ab|cd', cut)
        if fmt == "dense":
            sg = np.sort(np.abs(np.diag(S.data.cpu().numpy())))[::-1]
            rec_g = gtn.einsum('abx,xy,ycd->abcd', U, S, V).data.cpu().numpy()
        else:
            sg = np.sort(np.abs(np.diag(S.todense().data.cpu().numpy())))[::-1]
            rec_g = gtn.einsum('abx,xy,ycd->abcd', U, S, V).todense().data.cpu().numpy()
        sr = np.sort(np.abs(np.diag(Sr.data)))[::-1]
        n = min(len(sr), len(sg))
        assert np.abs(sr[:n] - sg[:n]).max() <= 1e-10 * sr[0], (kind, fmt)
        assert np.count_nonzero(sg > 1e-14 * sg[0]) == sum(counts)
        rec_r = O.einsum('abx,xy,ycd->abcd', Ur, Sr, Vr).data
        assert _relerr(rec_g, rec_r) < 1e-9, (kind, fmt)
    after = _ops.SVD_PATH_STATS
    if kind == "decaying":
        assert after["truncated"] > before["truncated"]


JOIN_CASES = [
    ((4, 2, 4, 8), (1, -1, -1, 1), '(ab)(cd)', 'standard', (1, -1)),
    ((4, 2, 4, 8), (1, -1, -1, 1), '(ab)(cd)', 'matrix', (-1, 1)),
    ((4, 2, 3, 4, 2), (1, -1, 0, -1, 1), '(abc)(de)', 'matrix', (-1, 1)),
    ((3, 4, 2, 5, 8), (0, 1, -1, 0, -1), '(ab)(cde)', 'standard', (1, -1)),
    ((4, 4, 2, 2), (1, 1, -1, -1), 'a(bc)d', 'matrix', (1, -1, -1)),
]


@pytest.mark.parametrize("case", range(len(JOIN_CASES)))
def test_join_split_legs_bit_exact(gtn, case):
    """user-level join_legs / split_legs (reference __init__.py:2947-3170), incl. hybrid '*' legs: the
    bosons-left step runs as a dense-mode sign+permute launch (per-element popcount signs)."""
    shape, stats, s, fmt, inter = JOIN_CASES[case]
    rng = np.random.RandomState(40 + case)
    a, A = _mk(gtn, shape, stats, rng)
    ref = O.join_legs(a, s, fmt, inter)
    got = A.join_legs(s, fmt, inter)
    assert [str(x) for x in got.statistics] == [str(x) for x in ref.statistics]
    assert (got.format, got.encoder) == (ref.format, ref.encoder)
    assert np.array_equal(got.data.cpu().numpy(), ref.data)
    back_r = O.split_legs(ref, s, stats, shape, inter)
    back_g = got.split_legs(s, stats, shape, inter)
    assert (back_g.format, back_g.encoder) == (back_r.format, back_r.encoder)
    assert np.array_equal(back_g.data.cpu().numpy(), back_r.data)
    # round trip (docs joinsplit.rst: (A - split(join(A))).norm = 0.0), compared in the canonical encoder
    rt = back_g.force_format('standard').force_encoder('canonical')
    assert np.array_equal(rt.data.cpu().numpy(), a.data)


@pytest.mark.parametrize("dtype", ["complex", "float"])
def test_truncated_svd_wide_subspace(gtn, dtype):
    """subspaces wider than 80 rows (chi >= 96) orthonormalise through the global-scratch variant of the
    pivoted-Cholesky kernel; kept singular triplets must match a full LAPACK SVD (reference SortedSVD,
    __init__.py:3931-3951) to 1e-10 * s_0."""
    import torch
    from grassmanntn_b200 import _engine
    rng = np.random.RandomState(77)
    p, q, k = 640, 768, 60
    def rnd(*s):
        return rng.randn(*s) + 1j * rng.randn(*s) if dtype == "complex" else rng.randn(*s)
    U0, _ = np.linalg.qr(rnd(p, p))
    V0, _ = np.linalg.qr(rnd(q, q))
    s0 = np.exp(-0.12 * np.arange(p))
    M = (U0 * s0) @ V0[:, :p].conj().T
    mats = [torch.from_numpy(M).cuda(), torch.from_numpy(np.ascontiguousarray(M[:600, :700])).cuda()]
    out = _engine.truncated_svd_batch(mats, [k, k - 7])
    assert out is not None, "certificate failed on a geometric spectrum"
    for (U, s, Vh), Mx, kk in zip(out, mats, (k, k - 7)):
        assert U.shape[1] >= 2 * kk + 8 > 80
        sr = np.linalg.svd(Mx.cpu().numpy(), compute_uv=False)
        assert np.abs(s[:kk] - sr[:kk]).max() <= 1e-10 * sr[0]
        Uk, Vk = U[:, :kk].cpu().numpy(), Vh[:kk].cpu().numpy()
        R = Mx.cpu().numpy() @ Vk.conj().T - Uk * s[:kk]
        assert np.abs(R).max() <= 1e-10 * sr[0]
        assert np.abs(Uk.conj().T @ Uk - np.eye(kk)).max() <= 1e-10
