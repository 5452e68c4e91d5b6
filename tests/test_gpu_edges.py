"""GPU edge cases and size-independent properties of the hot path (through the C ABI).

Small cases are compared with the CPU oracle; cases at the bench sizes (chi = 32 ... 64 sector
matrices, D = 32 tensors) use properties that do not need the oracle to finish: involutions of the
sign+permute kernel (bit-exact), join/split and hconjugate round trips (bit-exact), linearity of the
contraction, Eckart-Young for the truncated SVD (reference SortedSVD keeps the LARGEST singular values,
__init__.py:3931-3951), Hermitian eig reconstruction."""
import numpy as np
import pytest

import gtn_oracle as O

pytestmark = pytest.mark.gpu


def _mk(gtn, shape, stats, rng, cplx=True, trim=True):
    o = O.random_dense(shape, stats, dtype=complex if cplx else float, rng=rng, skip_trimming=not trim)
    return o, gtn.dense(o.data, statistics=stats)


def _np(x):
    return x.data.cpu().numpy()


# ---------------------------------------------------------------- small / ragged shapes vs the oracle
SMALL = [
    ('ab->ba', [((2, 2), (1, -1))]),
    ('ab->ba', [((1, 1), (1, -1))]),                       # one-element legs (index 0 only: even)
    ('a,a', [((4,), (1,)), ((4,), (-1,))]),                # vector . vector -> scalar
    ('ab,b->a', [((4, 8), (1, 1)), ((8,), (-1,))]),        # matrix . vector
    ('ij,jk->ik', [((3, 5), (0, 0)), ((5, 2), (0, 0))]),   # purely bosonic, ragged dims
    ('iaj,jbk->iabk', [((3, 2, 5), (0, 1, 0)), ((5, 4, 2), (0, -1, 0))]),   # fermions between bosons
    ('abcdefgh->hgfedcba', [((2,) * 8, (1, -1, 1, -1, 1, -1, 1, -1))]),     # 8 legs (maximum super-axes)
    ('abcdef,fedcba', [((2, 2, 2, 2, 2, 2), (1, 1, 1, -1, -1, -1)), ((2, 2, 2, 2, 2, 2), (1, 1, 1, -1, -1, -1))]),
]


@pytest.mark.parametrize("case", range(len(SMALL)))
def test_small_and_ragged_einsum(gtn, case):
    sub, ops = SMALL[case]
    rng = np.random.RandomState(500 + case)
    pairs = [_mk(gtn, s, st, rng, trim=False) for s, st in ops]
    ref = O.einsum(sub, *[p[0] for p in pairs])
    got = gtn.einsum(sub, *[p[1] for p in pairs])
    if isinstance(ref, O.Dense):
        assert got.shape == ref.shape and tuple(got.statistics) == tuple(ref.statistics)
        assert np.abs(_np(got) - ref.data).max() <= 1e-12 * max(np.abs(ref.data).max(), 1.0)
    else:
        assert abs(got - ref) <= 1e-12 * max(abs(ref), 1.0)


def test_einsum_errors(gtn):
    rng = np.random.RandomState(3)
    _, A = _mk(gtn, (4, 4), (1, -1), rng)
    _, B = _mk(gtn, (2, 4), (1, -1), rng)
    with pytest.raises(ValueError):
        gtn.einsum('ab,bc->ac', A, B)            # contracted dimensions differ
    with pytest.raises(ValueError):
        gtn.einsum('ab,bc->ac', A)               # operand count
    with pytest.raises(ValueError):
        gtn.einsum('abc->cba', A)                # leg count
    _, C = _mk(gtn, (4, 4), (-1, 1), rng)
    with pytest.raises(ValueError):
        gtn.einsum('ab,bc->ac', A, C)            # contraction of two non-conjugated legs (reference :1700-1740)


def test_eig_rejects_non_hermitian(gtn):
    rng = np.random.RandomState(4)
    _, A = _mk(gtn, (4, 4, 4, 4), (1, 1, -1, -1), rng)
    with pytest.raises(ValueError):
        A.eig('ab|cd')


def test_svd_cutoff_larger_than_rank_and_full(gtn):
    rng = np.random.RandomState(5)
    a, A = _mk(gtn, (4, 4, 4, 4), (1, 1, -1, -1), rng)
    for cut in (None, 64, 3):
        Ur, Sr, Vr = O.svd(a, 'ab|cd', cut)
        U, S, V = A.svd('ab|cd', cut)
        assert S.shape == Sr.shape
        sg = np.sort(np.abs(np.diag(_np(S))))[::-1]
        sr = np.sort(np.abs(np.diag(Sr.data)))[::-1]
        assert np.abs(sg - sr).max() <= 1e-10 * sr[0]
        if cut != 3:
            rec = gtn.einsum('abx,xy,ycd->abcd', U, S, V)
            assert np.abs(_np(rec) - a.data).max() <= 1e-12 * np.abs(a.data).max()


def test_svd_of_zero_sector_and_rank_one(gtn):
    """a tensor whose odd sector vanishes, and a rank-one tensor: rank rule s_i/(s_0+1e-14) > 1e-14
    (reference __init__.py:3939-3941) must give the oracle's bond dimension"""
    rng = np.random.RandomState(6)
    a, _ = _mk(gtn, (4, 4, 4, 4), (1, 1, -1, -1), rng)
    d = a.data.copy()
    par = np.array([bin(i).count('1') & 1 for i in range(4)])
    odd_rows = (par[:, None] ^ par[None, :]).astype(bool)
    d[odd_rows] = 0.0                                       # kill every element with odd (a,b) parity
    v = rng.rand(4, 4) + 1j * rng.rand(4, 4)
    v[odd_rows] = 0.0
    r1 = np.einsum('ab,cd->abcd', v, v.conj())
    for data in (d, r1):
        o = O.Dense(data, (1, 1, -1, -1))
        A = gtn.dense(data, statistics=(1, 1, -1, -1))
        Ur, Sr, Vr = O.svd(o, 'ab|cd', 16)         # int(16/2) = 8 per sector = the even sector's full rank
        U, S, V = A.svd('ab|cd', 16)
        assert S.shape == Sr.shape
        sg = np.sort(np.abs(np.diag(_np(S))))[::-1]
        sr = np.sort(np.abs(np.diag(Sr.data)))[::-1]
        assert np.abs(sg - sr).max() <= 1e-10 * max(sr[0], 1e-300)
        rec = gtn.einsum('abx,xy,ycd->abcd', U, S, V)
        assert np.abs(_np(rec) - data).max() <= 1e-10 * np.abs(data).max()


def test_block_non_power_of_two(gtn):
    """block tensors with ragged even/odd extents (reference block format, __init__.py:242-344)"""
    np.random.seed(7)
    R0 = gtn.random_block((5, 6, 3, 4), (1, 1, -1, -1), dtype=complex)     # not trimmed (like the reference)
    with pytest.raises(ValueError):
        R0.svd('ij|kl')                                                    # Error[BlockSVD]: not Grassmann-even
    A = gtn.trim_grassmann_odd(R0)
    even = [b for b in np.ndindex(2, 2, 2, 2) if sum(b) % 2 == 0]
    # transposition is an involution and preserves the norm
    B = gtn.einsum('ijkl->lkji', A)
    C = gtn.einsum('lkji->ijkl', B)
    assert abs(B.norm - A.norm) <= 1e-13 * A.norm
    for blk in even:
        assert np.array_equal(C.data[blk].cpu().numpy(), A.data[blk].cpu().numpy())
    # svd without cutoff reconstructs the tensor
    U, S, V = A.svd('ij|kl')
    R = gtn.einsum('ijx,xy,ykl->ijkl', U, S, V)
    for blk in even:
        x, y = R.data[blk].cpu().numpy(), A.data[blk].cpu().numpy()
        assert np.abs(x - y).max() <= 1e-12 * max(A.norm, 1.0)


# ---------------------------------------------------------------- properties at the bench sizes
def test_permute_involution_bit_exact_D32(gtn):
    rng = np.random.RandomState(8)
    _, A = _mk(gtn, (32, 32, 32, 32), (1, 1, -1, -1), rng, trim=False)
    for fwd, back in (('ijkl->jkli', 'jkli->ijkl'), ('ijkl->lkji', 'lkji->ijkl'), ('ijkl->kilj', 'kilj->ijkl')):
        B = gtn.einsum(fwd, A)
        C = gtn.einsum(back, B)
        assert np.array_equal(_np(C), _np(A)), fwd
        assert not np.array_equal(_np(B).ravel(), _np(A).ravel())


def test_switch_round_trips_bit_exact_D32(gtn):
    rng = np.random.RandomState(9)
    _, A = _mk(gtn, (32, 32, 32, 32), (1, -1, -1, 1), rng, trim=False)
    B = A.switch_format().switch_encoder().switch_encoder().switch_format()
    assert np.array_equal(_np(B), _np(A))
    H = A.hconjugate('ij|kl').hconjugate('ij|kl')
    assert np.array_equal(_np(H), _np(A))
    J = A.join_legs('(ij)(kl)', 'matrix', (1, -1))
    K = J.split_legs('(ij)(kl)', (1, -1, -1, 1), (32, 32, 32, 32), (1, -1)).force_format('standard').force_encoder('canonical')
    assert np.array_equal(_np(K), _np(A))


def test_contraction_linearity_D32(gtn):
    """einsum is bilinear: (A1 + 2 A2) . B == A1 . B + 2 A2 . B  to rounding, at D = chi = 32"""
    rng = np.random.RandomState(10)
    _, A1 = _mk(gtn, (32, 16, 16, 32), (1, 1, -1, 1), rng)
    _, A2 = _mk(gtn, (32, 16, 16, 32), (1, 1, -1, 1), rng)
    _, B = _mk(gtn, (32, 16, 16, 32), (-1, 1, -1, 1), rng)
    lhs = gtn.einsum('lxzk,jzxi->ijkl', A1 + A2 * 2.0, B)
    rhs = gtn.einsum('lxzk,jzxi->ijkl', A1, B) + gtn.einsum('lxzk,jzxi->ijkl', A2, B) * 2.0
    assert np.abs(_np(lhs) - _np(rhs)).max() <= 1e-12 * np.abs(_np(rhs)).max()


@pytest.mark.parametrize("chi,fmt", [(32, "dense"), (32, "block"), (64, "block"), (25, "block")])
def test_truncated_svd_eckart_young(gtn, chi, fmt):
    """|| T - U S V ||_F^2 == sum of the discarded s_i^2 of the sector matrices, and the kept values are
    the largest ones of each sector (full LAPACK SVD of the two sector matrices as the yard-stick)"""
    rng = np.random.RandomState(12)
    D = 32
    a = O.random_dense((D, D, D, D), (1, 1, -1, -1), dtype=complex, rng=rng)
    # graded spectrum (physical tensors decay): damp the tensor along a diagonal direction
    w = np.exp(-0.25 * np.arange(D))
    data = a.data * w[:, None, None, None] * w[None, :, None, None] * w[None, None, :, None] * w[None, None, None, :]
    A = gtn.dense(data, statistics=(1, 1, -1, -1))
    X = A if fmt == "dense" else A.toblock()
    U, S, V = X.svd('ab|cd', chi)
    R = gtn.einsum('abx,xy,ycd->abcd', U, S, V)
    Rd = _np(R if fmt == "dense" else R.todense())
    err2 = float(np.sum(np.abs(Rd - data) ** 2))
    # sector matrices in the parity-preserving order: rows (a,b) with p(a)+p(b) even / odd
    par = np.array([bin(i).count('1') & 1 for i in range(D)])
    rp = (par[:, None] ^ par[None, :]).ravel()
    M = data.reshape(D * D, D * D)
    disc = 0.0
    kE = chi // 2 if fmt == "dense" else -(-chi // 2)
    kO = chi // 2
    for sec, k in ((0, kE), (1, kO)):
        sub = M[np.ix_(rp == sec, rp == sec)]
        s = np.linalg.svd(sub, compute_uv=False)
        disc += float(np.sum(s[k:] ** 2))
    tot = float(np.sum(np.abs(data) ** 2))
    assert abs(err2 - disc) <= 1e-9 * tot, (err2, disc)


def test_eig_reconstruction_D16(gtn):
    rng = np.random.RandomState(13)
    a, A = _mk(gtn, (16, 16, 16, 16), (1, 1, -1, -1), rng)
    Ah = A.hconjugate('ab|cd')
    Hm = gtn.einsum('abxy,xycd->abcd', A, Ah)            # Hermitian positive
    U, L, V = Hm.eig('ab|cd')
    R = gtn.einsum('abx,xy,ycd->abcd', U, L, V)
    assert np.abs(_np(R) - _np(Hm)).max() <= 1e-10 * np.abs(_np(Hm)).max()


def test_truncated_eig_vs_oracle_D16(gtn):
    """eig with a cut far below the sector size (128 x 128 sectors, 8 eigenpairs each) goes through the truncated
    solver (subspace iteration + certificate): eigenvalues against the oracle's LAPACK path to 1e-10, and the kept
    part reproduces the Hermitian tensor's best approximation"""
    from grassmanntn_b200 import _ops
    rng = np.random.RandomState(17)
    # a decaying Hermitian tensor: H = A diag A^H built from a random tensor and a geometric spectrum in its bond
    a, A = _mk(gtn, (16, 16, 16), (1, 1, -1), rng)
    w = 0.8 ** np.arange(16)
    Wd = O.Dense(np.diag(w).astype(complex), (1, -1))
    Wg = gtn.dense(np.diag(w).astype(complex), statistics=(1, -1))
    h = O.einsum('abx,xy->aby', a, Wd)
    H = gtn.einsum('abx,xy->aby', A, Wg)
    hh = O.einsum('abx,xcd->abcd', h, O.hconjugate(h, 'ab|x'))
    HH = gtn.einsum('abx,xcd->abcd', H, H.hconjugate('ab|x'))
    before = dict(_ops.SVD_PATH_STATS)
    U, L, V = HH.eig('ab|cd', 16)
    assert _ops.SVD_PATH_STATS["truncated"] > before["truncated"]
    u, l, v = O.eig(hh, 'ab|cd', 16)
    lg, lr = np.sort(np.abs(np.diag(_np(L))))[::-1], np.sort(np.abs(np.diag(l.data)))[::-1]
    assert L.shape == l.shape
    assert np.abs(lg - lr).max() <= 1e-10 * lr[0]
    R = gtn.einsum('abx,xy,ycd->abcd', U, L, V)
    r = O.einsum('abx,xy,ycd->abcd', u, l, v)
    assert np.abs(_np(R) - r.data).max() <= 1e-9 * np.abs(r.data).max()


def test_rank_certificate_bounds_and_null_band_refinement(gtn):
    """_engine.deflated_norm_bound (the norm bound of M - U_D U_D^H M that decides the reference's rank rule for
    numerically rank-deficient sectors) against numpy: ||N||_2 <= bound <= ||N||_F, the Schatten-4 branch taken when the
    Frobenius norm alone is not conclusive; and refine_null_band lowering noise-level values of a full-SVD result."""
    import torch
    from grassmanntn_b200 import _engine as E
    rng = np.random.RandomState(23)
    p, q, r = 384, 512, 12
    A = rng.randn(p, r) + 1j * rng.randn(p, r)
    B = rng.randn(r, q) + 1j * rng.randn(r, q)
    noise = 1e-9 * (rng.randn(p, q) + 1j * rng.randn(p, q))
    M = A @ B + noise
    Ur, sr, Vr = np.linalg.svd(M, full_matrices=False)
    Nn = M - Ur[:, :r] @ (Ur[:, :r].conj().T @ M)
    two, fro = np.linalg.norm(Nn, 2), np.linalg.norm(Nn)
    Md = torch.from_numpy(M).cuda()
    U = torch.from_numpy(np.ascontiguousarray(Ur)).cuda()
    UhD = torch.from_numpy(np.ascontiguousarray(Ur[:, :r].conj().T)).cuda()
    thr = 0.5 * (two * 1.3 + fro)                      # Frobenius norm above, Schatten-4 bound below
    b1 = E.deflated_norm_bound(Md, UhD, U, U.shape[1], r, thr=fro * 2)        # conclusive at once: the Frobenius norm
    assert abs(b1 - fro) <= 1e-6 * fro
    b2 = E.deflated_norm_bound(Md, UhD, U, U.shape[1], r, thr=thr)
    assert two * (1 - 1e-9) <= b2 < fro, (two, b2, fro)
    assert b2 <= thr, (b2, thr, "the Schatten-4 bound of a flat noise spectrum is ~ n^(1/4) x the 2-norm, far below n^(1/2)")
    # full-SVD result with two noise values lifted above the rank rule's threshold: refinement certifies rank r
    M0 = torch.from_numpy(A @ B).cuda()
    (Uf, sf, Vf), = E.batched_svd([M0.clone()])
    s_fake = sf.copy()
    s_fake[r: r + 2] = 3e-14 * sf[0]
    # (left vectors additionally polluted at the 1e-12 level, ten times the Jacobi kernel's backward error)
    gen = torch.Generator(device="cpu"); gen.manual_seed(5)
    Uf = Uf + 1e-12 * torch.view_as_complex(torch.randn(tuple(Uf.shape) + (2,), generator=gen, dtype=torch.float64)).to(Uf.device)
    calls = dict(E.RANK_CHECK_STATS)
    _, s_ref, _ = E.refine_null_band(M0, (Uf, s_fake, Vf))
    assert E.RANK_CHECK_STATS["certified"] == calls["certified"] + 1
    assert int(np.sum(s_ref / (s_ref[0] + 1e-14) > 1e-14)) == r
    assert np.array_equal(s_ref[:r], sf[:r])


def test_chain_is_bitwise_reproducible(gtn):
    """Two executions of the same chain from the same engine state give the same BITS (round-1 review: the Jacobi
    rotation order depended on atomics timing and the norm on the order of atomicAdds).  Now: the dead-row threshold of
    the Jacobi kernels is double buffered by round parity, the sum of squares is completed in index order by the last
    CTA, split-K slices are added in index order."""
    import torch
    from grassmanntn_b200 import _engine as E, gauge2d as g
    def chain():
        for d in (E._trunc_iters_hint, E._trunc_rate, E._trunc_fail, E._trunc_probe, E._trunc_robust):
            d.clear()
        E.drop_graphs()
        T = g.zcap(g.load_initial_tensor()).toblock()
        norms = []
        for _ in range(5):
            T, n = g.trg(T, 32)[:2]
            norms.append(float(n))
        for _ in range(2):
            T, n = g.atrg2dy(T, T, 32)[:2]
            norms.append(float(n))
        return norms, T._bt.buf.clone()
    n1, b1 = chain()
    n2, b2 = chain()
    assert n1 == n2, (n1, n2)
    assert torch.equal(b1, b2)


def test_graph_replay_equals_eager_launches(gtn):
    """the steady-state truncated-SVD schedule replayed as a CUDA graph must give what the same launches
    give one by one (same kernels, same order): Tnorm of a TRG chain on the Z2 tensor, both ways"""
    from grassmanntn_b200 import _engine as E, gauge2d as g
    def chain(use_graphs):
        old = E.USE_GRAPHS
        E.USE_GRAPHS = use_graphs
        E._trunc_iters_hint.clear(); E._trunc_rate.clear(); E._trunc_fail.clear()
        try:
            T = g.zcap(g.load_initial_tensor()).toblock()
            out = []
            for _ in range(6):
                T, n = g.trg(T, 32)[:2]
                out.append(float(n))
            return out
        finally:
            E.USE_GRAPHS = old
    a, b = chain(False), chain(True)
    assert any(p.graphs for p in E._trunc_plans.values()), "no CUDA graph was captured"
    for x, y in zip(a[:3], b[:3]):
        assert abs(x - y) <= 1e-10 * abs(x), (a, b)          # the north-star tolerance on Tnorm
    for x, y in zip(a[3:], b[3:]):                   # from the 4th step on the cut goes through exact multiplets
        assert abs(x - y) <= 1e-6 * abs(x), (a, b)


def test_speculative_trg_step_and_misspeculation(gtn):
    """gauge2d.trg enqueues the rest of the step behind the replayed SVD graph and verifies the certificate
    at the end.  (a) same Tnorm as the non-speculative step; (b) a failed certificate (forced through the
    engine's test hook) is caught by verify(), the decomposition is resumed and the step repeated: results unchanged."""
    from grassmanntn_b200 import _engine as E, gauge2d as g
    def chain(spec, sabotage=False):
        old = g.SPECULATE
        g.SPECULATE = spec
        E._trunc_iters_hint.clear(); E._trunc_rate.clear(); E._trunc_fail.clear()
        try:
            T = g.zcap(g.load_initial_tensor()).toblock()
            out = []
            for i in range(8):
                # from the 4th step on every speculative run reports a failed certificate once
                E.FORCE_VERIFY_FAIL[0] = (lambda p: True) if (sabotage and i >= 3) else None
                T, n = g.trg(T, 32)[:2]
                out.append(float(n))
            return out, dict(g.SPEC_STATS)
        finally:
            g.SPECULATE = old
            E.FORCE_VERIFY_FAIL[0] = None
    g.STEP_GRAPH, old_graph = False, g.STEP_GRAPH             # the eager speculative path is under test here
    try:
        f0 = g.SPEC_STATS["failed"]
        (ref, _), (good, _), (bad, st) = chain(False), chain(True), chain(True, sabotage=True)
    finally:
        g.STEP_GRAPH = old_graph
    assert st["failed"] - f0 >= 1, st
    for other in (good, bad):
        for x, y in zip(ref[:3], other[:3]):
            assert abs(x - y) <= 1e-10 * abs(x), (ref, other)
        for x, y in zip(ref[3:], other[3:]):             # from the 4th step on the cut goes through exact multiplets
            assert abs(x - y) <= 1e-6 * abs(x), (ref, other)


@pytest.mark.gpu
def test_speculative_atrg_chain_and_misspeculation(gtn):
    """gauge2d.atrg2dy keeps its three dependent decompositions in flight behind each other (one workspace per
    call site) and verifies the certificates in order at the end of the step.  (a) same Tnorm as the
    non-speculative step; (b) a failed certificate at ONE stage (forced through the engine's test hook) is caught,
    that stage is resumed, the stages behind it are repeated; (c) same with every stage failing."""
    from grassmanntn_b200 import _engine as E, gauge2d as g
    T0 = g.zcap(g.load_initial_tensor()).toblock()
    for _ in range(2):
        T0, _ = g.trg(T0, 32)                                  # 32^4 site tensor: the truncated path is taken
    def chain(spec, sabotage=None):
        old = g.SPECULATE
        g.SPECULATE = spec
        E._trunc_iters_hint.clear(); E._trunc_rate.clear(); E._trunc_fail.clear()
        g.SPEC_STATS["speculated"] = g.SPEC_STATS["failed"] = 0
        try:
            T, out = T0, []
            for i in range(12):
                if sabotage is not None and i >= 4:
                    # the chosen stage(s) report a failed certificate on every speculative run
                    E.FORCE_VERIFY_FAIL[0] = lambda p: (isinstance(p.key[1], tuple) and p.key[1][0] == "atrg"
                                                        and (sabotage == "all" or p.key[1][2] == sabotage))
                fn = g.atrg2dx if i % 2 == 0 else g.atrg2dy
                T, n = fn(T, T, 32)[:2]
                out.append(float(n))
            return out, dict(g.SPEC_STATS)
        finally:
            g.SPECULATE = old
            E.FORCE_VERIFY_FAIL[0] = None
    g.STEP_GRAPH, old_graph = False, g.STEP_GRAPH             # the eager speculative chain is under test here
    try:
        ref, st0 = chain(False)
        runs = [chain(True), chain(True, sabotage=2), chain(True, sabotage="all")]
    finally:
        g.STEP_GRAPH = old_graph
    assert st0["speculated"] == 0
    assert runs[0][1]["speculated"] > 0, runs[0][1]
    assert runs[1][1]["failed"] >= 1 and runs[2][1]["failed"] >= 2, (runs[1][1], runs[2][1])
    for other, _ in runs:
        for i, (x, y) in enumerate(zip(ref, other)):
            # the cut goes through exact multiplets of the Z2 spectrum after the first steps
            tol = 1e-10 if i < 1 else 1e-6
            assert abs(x - y) <= tol * abs(x), (i, ref, other)


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["trg", "atrg"])
def test_whole_step_graph_matches_eager(gtn, algo):
    """In steady state a step is replayed as ONE CUDA graph (gauge2d._StepGraph) and verified from the read-backs.
    (a) the graph path engages; (b) same Tnorm chain as the eager speculative path; (c) a failed certificate at
    replay drops the graph and the step is redone eagerly: results unchanged."""
    from grassmanntn_b200 import _engine as E, gauge2d as g
    T0 = g.zcap(g.load_initial_tensor()).toblock()
    for _ in range(2):
        T0, _ = g.trg(T0, 32)
    nsteps = 14 if algo == "trg" else 24

    def step(T, i):
        if algo == "trg":
            return g.trg(T, 32)[:2]
        return (g.atrg2dx if i % 2 == 0 else g.atrg2dy)(T, T, 32)[:2]

    def chain(graph, sabotage=False):
        old, old_probe = g.STEP_GRAPH, E.PROBE_EVERY
        g.STEP_GRAPH = graph
        E.PROBE_EVERY = 0                # no periodic re-derivation of the counts: a recorded graph stays live
        E._trunc_iters_hint.clear(); E._trunc_rate.clear(); E._trunc_fail.clear(); E._trunc_probe.clear()
        g._step_graphs.clear(); g._steady.clear()
        g.STEP_GRAPH_STATS.pop("last_error", None)
        g.STEP_GRAPH_STATS.pop("capture_ms", None)
        for k in g.STEP_GRAPH_STATS:
            g.STEP_GRAPH_STATS[k] = 0
        try:
            T, out = T0, []
            for i in range(nsteps):
                # near the end one decomposition per step reports a failed certificate: the replayed graph is
                # dropped, the step is redone eagerly (speculative run fails too, is resumed)
                fails = [2]

                def hook(p):
                    fails[0] -= 1
                    return fails[0] >= 0
                E.FORCE_VERIFY_FAIL[0] = hook if (sabotage and i >= nsteps - 3) else None
                T, n = step(T, i)
                out.append(float(n))
            return out, dict(g.STEP_GRAPH_STATS)
        finally:
            g.STEP_GRAPH, E.PROBE_EVERY = old, old_probe
            E._trunc_probe.clear()       # (entries made with interval 0 must not outlive the test)
            E.FORCE_VERIFY_FAIL[0] = None
    ref, st0 = chain(False)
    assert st0["captured"] == 0 and st0["replayed"] == 0
    got, st1 = chain(True)
    assert st1["captured"] >= 1 and st1["replayed"] >= 2 and st1["failed_capture"] == 0, st1
    bad, st2 = chain(True, sabotage=True)
    assert st2["failed_certificate"] >= 1, st2
    for other in (got, bad):
        for i, (x, y) in enumerate(zip(ref, other)):
            tol = 1e-10 if i < 1 else 1e-6          # exact multiplets at the cut
            assert abs(x - y) <= tol * abs(x), (i, ref, other)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["complex", "float"])
def test_gemm_conj_transposed_b_operand(gtn, dtype):
    """GTN_GEMM_B_CONJ_TRANS: C = A * Bsrc^H read straight from the row-major N x K array (ragged sizes, split-K
    offsets, two groups in one launch) against torch, to rounding of a K-long dot product."""
    import torch
    from grassmanntn_b200 import _engine as E
    dt = torch.complex128 if dtype == "complex" else torch.float64
    gen = torch.Generator(device="cpu").manual_seed(3)
    def rnd(*shape):
        r = torch.randn(*shape, generator=gen, dtype=torch.float64)
        if dt == torch.complex128:
            r = torch.complex(r, torch.randn(*shape, generator=gen, dtype=torch.float64))
        return r.cuda()
    l1, K1, l2, n2, K2 = 48, 515, 13, 29, 70
    X, A2, B2 = rnd(l1, K1), rnd(l2, K2), rnd(n2, K2)
    buf = torch.cat([X.reshape(-1), A2.reshape(-1), B2.reshape(-1), torch.zeros(l1 * l1 + l2 * n2, dtype=dt, device="cuda")])
    oX, oA, oB = 0, X.numel(), X.numel() + A2.numel()
    oC1 = oB + B2.numel()
    oC2 = oC1 + l1 * l1
    k0 = 100                                   # second half of a split-K Gram matrix
    groups = [dict(a_off=oX + k0, b_off=oX + k0, c_off=oC1, lda=K1, ldb=K1, ldc=l1, m=l1, n=l1, k=K1 - k0, flags=1),
              dict(a_off=oA, b_off=oB, c_off=oC2, lda=K2, ldb=K2, ldc=n2, m=l2, n=n2, k=K2, alpha=-1.0, flags=1)]
    E.GemmPlan(groups, dt).run(buf, buf, buf)
    C1 = buf[oC1: oC1 + l1 * l1].view(l1, l1)
    C2 = buf[oC2: oC2 + l2 * n2].view(l2, n2)
    R1 = X[:, k0:] @ X[:, k0:].conj().T
    R2 = -(A2 @ B2.conj().T)
    assert float((C1 - R1).abs().max()) <= 1e-12 * float(R1.abs().max())
    assert float((C2 - R2).abs().max()) <= 1e-12 * float(R2.abs().max())


def test_dense_data_inplace_edit_is_seen(gtn):
    """the reference idiom `X.data[...] = v` / `X.data /= s` (gauge2d.py:1748): once `.data` has been handed out the
    array is the object's only storage, so later ops see in-place edits (round-1 advisor finding); a block made from
    the dense object does not alias it."""
    rng = np.random.RandomState(77)
    o, A = _mk(gtn, (4, 4), (1, -1), rng)
    tr0 = gtn.einsum('ii', A)
    Bk = A.toblock()
    A.data[0, 0] = 5.0
    d = o.data.copy()
    d[0, 0] = 5.0
    ref = O.einsum('ii', O.Dense(d, (1, -1)))
    assert abs(gtn.einsum('ii', A) - ref) <= 1e-13 * max(abs(ref), 1.0)
    assert abs(gtn.einsum('ii', Bk) - tr0) <= 1e-13 * max(abs(tr0), 1.0)
    X = gtn.einsum('ij->ji', A)
    view = X.data                      # a result keeps its block form until .data is asked for
    view *= 2.0
    assert abs(gtn.einsum('ii', X) - 2 * gtn.einsum('ii', gtn.einsum('ij->ji', A))) <= 1e-12 * max(abs(ref), 1.0)


# ---------------------------------------------------------------- unwritten permutations (_ops.LazyPermute)
def _bits(gtn, x):
    """every stored element of a result, as numpy (forces an unwritten permutation to be written)"""
    if isinstance(x, (tuple, list)):
        return [_bits(gtn, y) for y in x]
    bt = x._bt if getattr(x, "_bt", None) is not None else x._get_bt()
    return (dict(bt.off), bt.buf.cpu().numpy().copy())


def _same(a, b):
    if isinstance(a, list):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a[0] == b[0] and a[1].shape == b[1].shape and np.array_equal(a[1], b[1])


def _lazy_chain(gtn, T, fmt):
    X = T.toblock() if fmt == "block" else T
    P1 = gtn.einsum('ijkl->jkli', X)                       # signed permutation
    P2 = gtn.einsum('abcd->dcab', P1)                      # of an unwritten one: refers to X
    H = P2.hconjugate('ab|cd')                             # conjugated + signed + permuted, still unwritten
    R1 = gtn.einsum('abcd,abef->cdef', H, P1)              # both operands packed from X
    R2 = gtn.einsum('abcd,dbef->feac', P2, X)              # GEMM result left in its natural leg order ...
    R3 = gtn.einsum('feac,face', R2, P1)                   # ... and packed from there (scalar)
    tr = gtn.einsum('ijij', P1)                            # trace over legs of an unwritten permutation
    U, S, V = P1.svd('ab|cd')                              # matricised from X
    U2, S2, V2 = H.svd('ab|cd', 5)
    HH = H.hconjugate('ab|cd')                             # conjugate of an unwritten conjugate
    n1 = P2.norm
    Sc = P2 * 0.5
    return [P1, P2, H, R1, R2, U, S, V, U2, S2, V2, HH, Sc], [R3, tr, n1]


@pytest.mark.parametrize("fmt", ["dense", "block"])
@pytest.mark.parametrize("cplx", [True, False])
def test_unwritten_permutations_equal_written_ones(gtn, fmt, cplx):
    """a chain of leg permutations, conjugates, contractions, traces and decompositions with the intermediates left
    unwritten (consumers read the stored source through composed sign tables) against the same chain with every
    intermediate written: identical bits everywhere (signs are exact and each element is moved, never recomputed);
    only norms may differ in the last place (summation order)."""
    from grassmanntn_b200 import _ops
    rng = np.random.RandomState(91)
    _, T = _mk(gtn, (4, 8, 4, 8), (1, 1, -1, -1), rng, cplx=cplx)
    saved = _ops.LAZY_PERMUTE
    try:
        _ops.LAZY_PERMUTE = False
        ten0, sc0 = _lazy_chain(gtn, T, fmt)
        ten0 = _bits(gtn, ten0)
        _ops.LAZY_PERMUTE = True
        for k in _ops.LAZY_STATS:
            _ops.LAZY_STATS[k] = 0
        ten1, sc1 = _lazy_chain(gtn, T, fmt)
        st = dict(_ops.LAZY_STATS)
        unwritten = [getattr(x, "_bt", None) is not None and x._bt.pending() is not None for x in ten1]
        ten1 = _bits(gtn, ten1)
    finally:
        _ops.LAZY_PERMUTE = saved
    assert st["created"] >= 4 and st["materialised"] <= 1, st     # (the scaled copy Sc is written when it is scaled)
    assert unwritten[:3] == [True, True, True] and unwritten[11], unwritten
    for k, (a, b) in enumerate(zip(ten0, ten1)):
        assert _same(a, b), k
    for a, b in zip(sc0, sc1):
        assert abs(a - b) <= 4e-16 * max(abs(a), 1.0)


def test_unwritten_permutation_is_not_an_alias(gtn):
    """einsum returns a new tensor in the reference: writing into the source's blocks afterwards (`.data` views) must
    not change a permutation of it that has not been written yet"""
    rng = np.random.RandomState(92)
    _, T = _mk(gtn, (4, 4, 4, 4), (1, 1, -1, -1), rng)
    B = T.toblock()
    P = gtn.einsum('ijkl->lkji', B)
    C = gtn.einsum('ijkl->lkji', B.copy())
    want = _bits(gtn, C)
    views = B.data
    for idx in np.ndindex(views.shape):
        views[idx] *= 0.0
    assert float(B.norm) == 0.0
    assert _same(_bits(gtn, P), want)


@pytest.mark.parametrize("algo", ["trg", "atrg", "hotrg3dz"])
def test_steps_with_unwritten_permutations_equal_written_ones(gtn, algo):
    """whole coarse-graining steps on the Z2 tensor with and without the unwritten-permutation path: same Tnorm and
    the same tensor to rounding of the norm (the data path is bit-identical; the norm of an unwritten result is summed
    over its source, i.e. in another order)"""
    from grassmanntn_b200 import _ops, gauge2d as g
    old_graph, old_spec = g.STEP_GRAPH, g.SPECULATE
    g.STEP_GRAPH = False

    def run():
        if algo == "hotrg3dz":
            T6 = g.load_initial_tensor()
            T, n = g.hotrg3dz(T6, T6, 8)
            return [n], T
        T = g.zcap(g.load_initial_tensor()).toblock()
        ns = []
        for i in range(3):
            if algo == "trg":
                T, n = g.trg(T, 8)
            else:
                T, n = (g.atrg2dy if i % 2 == 0 else g.atrg2dx)(T, T, 8)
            ns.append(n)
        return ns, T
    saved = _ops.LAZY_PERMUTE
    try:
        _ops.LAZY_PERMUTE = False
        n0, T0 = run()
        _ops.LAZY_PERMUTE = True
        n1, T1 = run()
    finally:
        _ops.LAZY_PERMUTE = saved
        g.STEP_GRAPH, g.SPECULATE = old_graph, old_spec
    for a, b in zip(n0, n1):
        assert abs(a - b) <= 1e-15 * abs(a), (n0, n1)
    (off0, b0), (off1, b1) = _bits(gtn, T0), _bits(gtn, T1)
    assert off0 == off1
    assert np.abs(b0 - b1).max() <= 1e-14 * np.abs(b0).max()
