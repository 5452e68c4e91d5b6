"""Coarse-graining drivers of the 2D gauge-theory example on the B200 ops.

Same call signatures and return values as the reference's gauge2d.trg / atrg2dy / atrg2dx / zcap /
logZ (reference gauge2d.py:1591-1889, and the block twin gauge2d_block.py); one implementation
serves dense and block tensors because every op underneath works on the parity-blocked device
storage.  The tensor-network contractions are the algorithm (Levin-Nave TRG; ATRG) and therefore
use the same index patterns as the reference; everything else (no deep copies, no progress bars,
normalisation fused into one scale kernel) is new.

The initial-tensor construction (tensor_preparation, get_ABtensors, compress_*) is out of scope
(SURVEY.md section 8): load the fixture with `load_initial_tensor`.
"""
import math
import os

import numpy as np

import grassmanntn_b200 as gtn

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_initial_tensor(path=None):
    """Z2, N_f=1, beta=m=q=a=1, mu=0 site tensor produced once by the reference
    (tests/golden/make_z2_tensor.py): shape (8,8,8,8,2,2), statistics (1,1,-1,-1,0,0)."""
    path = path or os.path.join(_GOLDEN, "z2_initial_tensor.npz")
    z = np.load(path)
    stats = tuple(int(s) for s in z["statistics"])
    return gtn.dense(z["data"], statistics=stats, encoder=str(z["encoder"]), format=str(z["format"]))


def load_AB_tensors(path=None):
    """The site tensor A[i,j,k,l] (bosonic) and the link tensor B[I,J,K,L,i,j,k,l] of the same model as produced by the
    reference's get_ABtensors (symbolic Berezin integration, gauge2d.py:139-257; tests/golden/make_z2_prep_golden.py):
    the inputs of tensor_from_AB."""
    z = np.load(path or os.path.join(_GOLDEN, "z2_prep.npz"))
    B = np.zeros(tuple(int(x) for x in z["B_shape"]), dtype=complex)
    B[tuple(z["B_coords"].T.astype(np.int64))] = z["B_vals"]
    return (gtn.dense(np.asarray(z["A"]).astype(complex), statistics=tuple(int(s) for s in z["A_stats"])),
            gtn.dense(B, statistics=tuple(int(s) for s in z["B_stats"])))


SPECULATE = bool(int(os.environ.get("GTN_SPECULATE", "1")))
SPEC_STATS = {"speculated": 0, "failed": 0}       # decompositions run speculatively / whose certificate then failed


def _normalised(T):
    """(T / |T|, |T|) for a step's own un-normalised result: scaled in place (T is not used again by the callers);
    a result that is still an unwritten permutation (_ops.LazyPermute) is written once, scaled on the way"""
    Tnorm = T.norm
    if getattr(T, "_bt", None) is None or getattr(T, "_data", None) is not None:
        return T * (1.0 / Tnorm), Tnorm
    T._bt.scale_(1.0 / Tnorm)
    return T, Tnorm


def zcap(T):
    """sum over the two bosonic legs (reference gauge2d.py:1591-1597)."""
    n = T.shape[4]
    capper = gtn.dense(np.full((n, n), 1.0), statistics=(0, 0))
    if isinstance(T, gtn.block):
        capper = capper.toblock()
    return gtn.einsum("IJKLij,ij->IJKL", T, capper)


def logZ(T, boundary_conditions="periodic"):
    """log of the trace of T with (anti-)periodic boundary in the second direction
    (reference gauge2d.py:1599-1615).  Like gauge2d_block.py:1608 the anti-periodic sign is NOT
    applied to block tensors."""
    if boundary_conditions == "anti-periodic" and not isinstance(T, gtn.block):
        T = _flip_leg_parity(T, 1)
    Z = gtn.einsum("IJIJ", T)
    return np.log(Z)


def _flip_leg_parity(T, leg):
    """(-1)^{p} on one leg (the anti-periodic boundary sign, reference gauge2d.py:1606-1611): a
    block-constant sign, applied with the scale kernel to the odd blocks of that leg."""
    from grassmanntn_b200 import _cabi, _engine
    bt = T._get_bt().clone()
    for p in bt.live():
        if dict(zip(bt.faxes, p)).get(leg, 0) == 1:
            v = bt.buf[bt.off[p]: bt.off[p] + bt.block_size(p)]
            _cabi.check(_cabi.lib.gtn_scale(_engine._ptr(v), v.numel(), _engine.dtype_code(v.dtype), -1.0, 0.0,
                                            _engine._stream()), "gtn_scale")
            _cabi.count()
    return gtn.dense._from_bt(bt, T.encoder)


# ------------------------------------------------------------------------------------------------
#  whole-step CUDA graphs
# ------------------------------------------------------------------------------------------------
STEP_GRAPH = bool(int(os.environ.get("GTN_STEP_GRAPH", "1")))
STEP_GRAPH_MAX_BYTES = 160 << 20          # site tensors up to chi = 64 (the graph's pool keeps a step's intermediates)
STEP_GRAPH_STATS = {"captured": 0, "replayed": 0, "failed_capture": 0, "failed_certificate": 0}
_step_graphs = {}                         # key -> _StepGraph, or False after a failed capture
_capture_failures = {}
_POOL = [None, None, None]       # pool handle, keeper graph, keeper tensor
_steady = {}                              # key -> (consecutive verified speculative eager steps, hints then)


def _drop_step_graphs():
    _step_graphs.clear()
    _steady.clear()


def _register_dropper():
    from . import _engine as E
    if _drop_step_graphs not in E._graph_droppers:
        E._graph_droppers.append(_drop_step_graphs)


_register_dropper()


def freeze(flag=True):
    """Freeze (or release) the engine's adaptation: the truncated SVD keeps its learnt iteration counts (no lowering,
    no periodic re-derivation) and recorded step graphs are not recorded again.  For timing loops and for
    production runs that repeat the same step shape; certificates are still verified on every step and a
    failed one still falls back to the eager path."""
    from . import _engine as E
    E.ADAPT[0] = not flag


def _bt_of(T):
    return T._bt if isinstance(T, gtn.block) else T._get_bt()


def _wrap_bt(bt, like):
    return gtn.block._from_bt(bt, like.shape) if isinstance(like, gtn.block) else gtn.dense._from_bt(bt, like.encoder)


def _step_key(name, T, *params):
    return (name, type(T).__name__, getattr(T, "encoder", None), _bt_of(T).layout_key()) + params


class _StepGraph:
    """One coarse-graining step, from the site tensor to the un-normalised result and its squared norm, recorded
    as ONE CUDA graph: permutes, packs, the truncated SVD schedules (enqueued inline), unpacks, contractions.
    Bond dimensions are data dependent in general (rank rule), so the graph encodes the steady state 'every
    sector keeps its full cut'; after every replay the certificates of its decompositions and that assumption
    are verified from the pinned read-backs (one synchronisation per step), and the caller falls back to the
    eager path when they fail.  Normalisation (one scale launch) happens outside: the norm is known to the host
    only after the synchronisation."""

    def __init__(self, T, body):
        import torch
        from . import _cabi, _engine as E
        _bt_of(T).buf                                 # (an unwritten permutation is written: the graph needs an input
        self.static = _bt_of(T).clone()               #  buffer of its own, not a view of this particular tensor)
        assert self.static.pending() is None
        self.key = self.static.key()
        Tin = _wrap_bt(self.static, T)
        self.norm_host = torch.zeros(1, dtype=torch.float64).pin_memory()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        c0 = _cabi.launch_count
        g = torch.cuda.CUDAGraph()
        # one memory pool for all step graphs: a graph recorded again (shorter SVD schedules) reuses the blocks of the
        # one it replaces instead of paying ~100 cudaMallocs (60-600 ms measured for an ATRG step).  Safe because
        # step graphs never run concurrently and every result is copied out of the pool right after its replay
        if _POOL[0] is None:
            # torch frees a pool with its last graph: a one-node keeper graph holds it for the life of the process
            handle = torch.cuda.graph_pool_handle()
            keeper = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                keeper.capture_begin(pool=handle)
                try:
                    _POOL[2] = torch.zeros(1, device="cuda")
                finally:
                    keeper.capture_end()
            _POOL[0], _POOL[1] = handle, keeper
        E.CAPTURING_STEP[0] = True
        try:
            with torch.cuda.stream(side):
                g.capture_begin(pool=_POOL[0])
                try:
                    out, pend, norm_of = body(Tin)
                    self.obt = _bt_of(out)
                    self.acc = _bt_of(norm_of).sumsq()
                    self.norm_host.copy_(self.acc, non_blocking=True)
                finally:
                    g.capture_end()
        finally:
            E.CAPTURING_STEP[0] = False
        cur.wait_stream(side)
        self.launches = _cabi.launch_count - c0
        _cabi.launch_count = c0                       # captured, not executed
        self.g, self.like = g, out
        self.pend = [p for p in pend if p is not None]
        self.its = [p.it for p in self.pend]
        self.replays = 0

    def replay(self, T):
        from . import _cabi
        bt = _bt_of(T)
        if bt.layout_key() != self.key or bt.buf.numel() != self.static.buf.numel():
            return None
        self.static.buf.copy_(bt.buf)
        self.g.replay()
        _cabi.count(self.launches)
        self.replays += 1
        for p in self.pend:
            SPEC_STATS["speculated"] += 1
            if not p.reverify():                      # the first one synchronises the stream
                SPEC_STATS["failed"] += 1
                return None
        Tnorm = math.sqrt(float(self.norm_host[0]))
        new = self.obt.clone()
        new.scale_(1.0 / Tnorm)
        return _wrap_bt(new, self.like), Tnorm

    def stale(self):
        """a decomposition now needs fewer iterations than were recorded, or its count is being derived afresh"""
        from . import _engine as E
        hints = E._trunc_iters_hint
        if not E.ADAPT[0]:
            return False
        return any(p.key not in hints or hints[p.key] < it for p, it in zip(self.pend, self.its))


def _graph_step(key, T, body):
    """the step through its recorded graph, or None (not in steady state / capture impossible / certificate failed)"""
    import torch
    from . import _engine as E
    sg = _step_graphs.get(key)
    if sg is False or E.PROF.enabled:
        return None
    if sg is None:
        n, _ = _steady.get(key, (0, None))
        bt = _bt_of(T)
        if n < 2 or bt.stored_elems() * (16 if bt.dtype.is_complex else 8) > STEP_GRAPH_MAX_BYTES:
            return None
        import time as _time
        t0 = _time.perf_counter()
        try:
            sg = _StepGraph(T, body)
        except Exception as exc:                      # NotCapturable, or CUDA refusing an operation during capture
            STEP_GRAPH_STATS["failed_capture"] += 1
            STEP_GRAPH_STATS["last_error"] = repr(exc)[:300]
            torch.cuda.synchronize()
            _steady[key] = (0, None)                  # a few eager steps first (they build what was missing)
            _capture_failures[key] = _capture_failures.get(key, 0) + (0 if isinstance(exc, E.NotCapturable) else 1)
            if _capture_failures[key] >= 3:
                _step_graphs[key] = False             # CUDA keeps refusing: stay eager on this layout
            return None
        STEP_GRAPH_STATS["captured"] += 1
        STEP_GRAPH_STATS["capture_ms"] = (STEP_GRAPH_STATS.get("capture_ms", []) + [round((_time.perf_counter() - t0) * 1e3, 1)])[-16:]
        if len(_step_graphs) >= 8:
            _step_graphs.pop(next(iter(_step_graphs)))
        _step_graphs[key] = sg
    r = sg.replay(T)
    if r is None:
        STEP_GRAPH_STATS["failed_certificate"] += 1
        _step_graphs.pop(key, None)
        _steady[key] = (0, None)
        return None
    STEP_GRAPH_STATS["replayed"] += 1
    if sg.stale():
        _step_graphs.pop(key, None)                   # record it again with the shorter schedules ...
        _steady[key] = (0, None)                      # ... after eager steps have run (and built) them
    return r


def _note_eager(key, pendings):
    """bookkeeping after an eager step: it counts towards the steady state when every decomposition was
    speculated and verified and the iteration hints did not move"""
    from . import _engine as E
    ok = bool(pendings) and all(p is not None and p.ok for p in pendings)
    hints = tuple(E._trunc_iters_hint.get(p.key) for p in pendings) if ok else None
    n, last = _steady.get(key, (0, None))
    _steady[key] = ((n + 1) if (ok and (last is None or last == hints)) else (1 if ok else 0), hints)


def _trg_enqueue(T, dcut, resume=None, error_test=False):
    """enqueue one TRG step; returns (un-normalised T', pending decomposition or None, trace error or None)"""
    T1 = gtn.einsum("ijkl->jkli", T)
    T2 = gtn.einsum("ijkl->klij", T)
    pending = None
    if resume is not None:
        res = gtn.svd_many([T1, T2], "ab|cd", dcut, resume=resume, site=("trg",))
    elif SPECULATE:
        # steady state: the truncated SVD replays its CUDA graph and the rest of the step is enqueued behind it
        # without waiting for the certificate (one synchronisation per step); verified by the caller
        res, pending = gtn.svd_many([T1, T2], "ab|cd", dcut, speculative=True, site=("trg",))
    else:
        res = gtn.svd_many([T1, T2], "ab|cd", dcut, site=("trg",))
    (U1, S1, V1), (U2, S2, V2) = res
    sq = gtn.sqrt(S1)
    U1 = gtn.einsum("abx,xc->abc", U1, sq)
    V1 = gtn.einsum("ax,xbc->abc", sq, V1)
    sq = gtn.sqrt(S2)
    U2 = gtn.einsum("abx,xc->abc", U2, sq)
    V2 = gtn.einsum("ax,xbc->abc", sq, V2)
    VV = gtn.einsum("kwz,lxw->lxzk", V1, V2)
    UU = gtn.einsum("yxi,zyj->jzxi", U1, U2)
    Tn = gtn.einsum("lxzk,jzxi->ijkl", VV, UU)
    err = None
    if error_test:
        Z1 = gtn.einsum("ijkl,klij", T, T)
        Z2 = gtn.einsum("ijij", Tn)
        err = np.abs(1 - Z2 / Z1)
    return Tn, pending, err


def trg(T, dcut=64, iternum=None, error_test=False):
    """One Levin-Nave TRG step (reference gauge2d.py:1647-1755).  T: shape (m,n,m,n), statistics
    (1,1,-1,-1).  Returns (T', Tnorm[, err]).
    Steady state (same layout as the last steps, every sector at its full cut): the whole step replays as one
    CUDA graph (_StepGraph); otherwise it is enqueued eagerly with a speculative truncated SVD."""
    if [T.shape[0], T.shape[1]] != [T.shape[2], T.shape[3]]:
        gtn.error("Error[trg]: The shape must be of the form (m,n,m,n)!")
    if gtn.make_list(T.statistics) != [1, 1, -1, -1]:
        gtn.error("Error[trg]: The statistics must be (1,1,-1,-1)!")
    use_graph = STEP_GRAPH and SPECULATE and not error_test
    if use_graph:
        key = _step_key("trg", T, dcut)

        def body(X):
            Tn, pending, _ = _trg_enqueue(X, dcut)
            return Tn, [pending], Tn
        r = _graph_step(key, T, body)
        if r is not None:
            return r
    Tn, pending, err = _trg_enqueue(T, dcut, error_test=error_test)
    Tn, Tnorm = _normalised(Tn)
    if pending is not None:
        SPEC_STATS["speculated"] += 1
        if not pending.verify():
            SPEC_STATS["failed"] += 1
            Tn, _, err = _trg_enqueue(T, dcut, resume=pending, error_test=error_test)
            Tn, Tnorm = _normalised(Tn)
    if use_graph:
        _note_eager(key, [pending])
    return (Tn, Tnorm, err) if error_test else (Tn, Tnorm)


def _svd_stage(objs, string, cut, site, resume, pend):
    """One decomposition of a multi-stage step.  Steady state: speculative (the replayed SVD graph is only
    enqueued; `pend[site]` holds the unverified run).  `resume`: continue a run whose certificate failed."""
    if resume is not None:
        return gtn.svd_many(objs, string, cut, resume=resume, site=site)
    if SPECULATE:
        res, pend[site] = gtn.svd_many(objs, string, cut, speculative=True, site=site)
        return res
    return gtn.svd_many(objs, string, cut, site=site)


def _atrg_pass(T1, T2, same, dcut, intermediate_dcut, sites, st, start, resume):
    """enqueue the stages >= start of one ATRG step (earlier stages: results kept in `st`); returns the
    un-normalised T' and {site: unverified decomposition}"""
    pend = {}
    if start <= 0:
        st["a"] = _svd_stage([T1] if same else [T1, T2], "li|jk", intermediate_dcut, sites[0],
                             resume if start == 0 else None, pend)
    (U1, S1, V1), (U2, S2, V2) = st["a"][0], st["a"][-1]
    if start <= 1:
        A = V1
        B = gtn.einsum("lia,ab->lib", U1, S1)
        C = gtn.einsum("ab,bjk->ajk", S2, V2)
        D = U2
        M = gtn.einsum("ajk,jib->aibk", C, B)
        st["b"] = _svd_stage([M], "ai|bk", intermediate_dcut, sites[1], resume if start == 1 else None, pend)
        st["AD"] = (A, D)
    A, D = st["AD"]
    U, S, V = st["b"][0]
    if start <= 2:
        sq = gtn.sqrt(S)
        Y = gtn.einsum("abx,xc->abc", U, sq)
        X = gtn.einsum("ax,xbc->abc", sq, V)
        Q1 = gtn.einsum("iax,xbj->ijab", D, Y)
        Q2 = gtn.einsum("kya,ylb->abkl", X, A)
        Q = gtn.einsum("ijab,abkl->ijkl", Q1, Q2)
        st["c"] = _svd_stage([Q], "ij|kl", dcut, sites[2], resume if start == 2 else None, pend)
    U, S, V = st["c"][0]
    sq = gtn.sqrt(S)
    H = gtn.einsum("abx,xc->abc", U, sq)
    G = gtn.einsum("ax,xbc->abc", sq, V)
    H = gtn.einsum("lai->ila", H)
    G = gtn.einsum("kaj->ajk", G)
    return gtn.einsum("ila,ajk->ijkl", H, G), pend


def atrg2dy(T1, T2, dcut=64, intermediate_dcut=None, iternum=None, error_test=False, alignment="y"):
    """One ATRG step along y (reference gauge2d.py:1761-1869).
    The three decompositions depend on each other; in steady state each replays its truncated-SVD graph without
    reading the certificate back, the next stage is enqueued behind it, and the step synchronises once (its
    norm).  The certificates are then verified in order; the first stage that fails is resumed from its
    workspace state and everything after it is repeated.  Once that has worked twice in a row on the same layout
    the whole step (T1 is T2) is recorded as one CUDA graph (_StepGraph)."""
    if intermediate_dcut is None:
        intermediate_dcut = dcut
    same = T1 is T2
    sites = [("atrg", alignment, i) for i in (1, 2, 3)]
    use_graph = STEP_GRAPH and SPECULATE and same and not error_test
    if use_graph:
        key = _step_key("atrg" + alignment, T1, dcut, intermediate_dcut)

        def body(X):
            Xr = gtn.einsum("ijkl->lijk", X)
            Tn, pend = _atrg_pass(Xr, Xr, True, dcut, intermediate_dcut, sites, {}, 0, None)
            return Tn, [pend.get(s_) for s_ in sites], Tn
        r = _graph_step(key, T1, body)
        if r is not None:
            return r
    T1o, T2o = T1, T2
    T1 = gtn.einsum("ijkl->lijk", T1)
    T2 = T1 if same else gtn.einsum("ijkl->lijk", T2)
    st, start, resume, clean = {}, 0, None, True
    while True:
        T, pend = _atrg_pass(T1, T2, same, dcut, intermediate_dcut, sites, st, start, resume)
        err = None
        if error_test:
            Z1 = gtn.einsum("IJIK,iKiJ", T1o, T2o)
            Z2 = gtn.einsum("IJIJ", T)
            err = np.abs(1 - Z2 / Z1)
        T, Tnorm = _normalised(T)
        # verify the speculated stages in order; stages behind a failed one worked on unverified input
        bad = None
        for i, site in enumerate(sites):
            p = pend.get(site)
            if p is None:
                continue
            SPEC_STATS["speculated"] += 1
            if bad is not None:
                p.discard()
            elif not p.verify():
                SPEC_STATS["failed"] += 1
                bad = i
        if bad is None:
            break
        start, resume, clean = bad, pend[sites[bad]], False
    if use_graph:
        _note_eager(key, [pend.get(s_) for s_ in sites] if clean else [])
    return (T, Tnorm, err) if error_test else (T, Tnorm)


def _swap_xy(T):
    return gtn.einsum("jikl->jilk", gtn.einsum("ijkl->jikl", T))


def atrg2dx(T1, T2, dcut=64, intermediate_dcut=None, iternum=None, error_test=False):
    """One ATRG step along x: swap the legs, atrg2dy, swap back (reference gauge2d.py:1871-1889)."""
    same = T1 is T2
    T1 = _swap_xy(T1)
    T2 = T1 if same else _swap_xy(T2)
    out = atrg2dy(T1, T2, dcut, intermediate_dcut, iternum, error_test, alignment="x")
    return (_swap_xy(out[0]),) + tuple(out[1:])


def hotrg3dz(T1, T2, dcut=64, intermediate_dcut=None, iternum=None, error_test=False):
    """Flavour-direction HOTRG-type step merging two 6-leg site tensors (legs 5,6 bosonic)
    (reference gauge2d.py:1891-2145).  Same contraction network as the reference: three
    decompositions per input, Gram matrices of the XX / YY environments, isometries from the
    truncated eigen-decompositions, then the final merge.  Returns (T, Tnorm[, err])."""
    E = gtn.einsum
    if intermediate_dcut is None:
        intermediate_dcut = dcut
    T1o, T2o = T1, T2

    def split3(T):
        T = E('i1 i2 i3 i4 mn-> i1 i3 mn i2 i4', T)
        X, S, Y = T.svd('(i1 i3 m)(n i2 i4)', intermediate_dcut)
        sq = gtn.sqrt(S)
        X = E('i1 i3 m a,ab->i1 i3 b m', X, sq)
        Y = E('ab,b n i2 i4->n a i2 i4', sq, Y)
        X, S, Pm = X.svd('(i1 i3)(a m)', intermediate_dcut)
        X = E('i1 i3 x, xy -> i1 i3 y', X, S)
        Q, S, Y = Y.svd('(na)(i2 i4)', intermediate_dcut)
        Y = E('xy, y i2 i4 -> x i2 i4', S, Y)
        return X, Pm, Q, Y

    same = T1 is T2
    X1, P1, Q1, Y1 = split3(T1)
    X2, P2, Q2, Y2 = (X1, P1, Q1, Y1) if same else split3(T2)

    def isometry(A1, A2, fused, h1, g1, h2, g2):
        # the reference forms the outer product and then moves the legs by four single swaps (gauge2d.py:1964-1968,
        # :2027-2031), each a full pass over the chi^2 D^4 tensor; Grassmann reorderings compose, so the product is
        # written straight into the final leg order (same signs, bit for bit: one launch instead of five)
        AA = E(fused, A1, A2)
        cA = AA.hconjugate(h1)
        Ma = E(g1, cA, AA)
        Ua, Sa, _ = Ma.eig('(I J)(i j)', dcut)
        cB = AA.hconjugate(h2)
        Mb = E(g2, AA, cB)
        Ub, Sb, _ = Mb.eig('(I J)(i j)', dcut)
        return AA, (Ua if Sa.shape[0] < Sb.shape[0] else Ub)

    XX, Ux = isometry(X1, X2, 'i1 i3 a, j1 j3 b -> i3 j3 ab i1 j1',
                      '(i3 j3 ab)(i1 j1)', ' I1 J1 i3 j3 ab, i3 j3 ab i1 j1 -> I1 J1 i1 j1',
                      '(i3 j3)(ab i1 j1)', ' I3 J3 ab i1 j1, ab i1 j1 i3 j3  -> I3 J3 i3 j3')
    cUx = Ux.hconjugate('ij|a')
    # (reference :1994-1998: two more swaps of XX, then the contraction, then a leg rotation of the result -- folded
    # into the index strings of the contractions themselves)
    Xp = E('s i3 j3, i3 j3 kl i1 j1 -> s kl j1 i1', cUx, XX)
    Xp = E('s kl j1 i1, i1 j1 t -> t s kl', Xp, Ux)
    Xp = E('t s kl , kam -> t s al m', Xp, P1)
    Xp = E('t s al m , lbm -> t s ab m', Xp, P2)

    YY, Uy = isometry(Y2, Y1, 'b j2 j4, a i2 i4 -> i4 j4 ba i2 j2',
                      '(i4 j4 ba)(i2 j2)', ' I2 J2 i4 j4 ba, i4 j4 ba i2 j2 -> I2 J2 i2 j2',
                      '(i4 j4)(ba i2 j2)', ' I4 J4 ba i2 j2, ba i2 j2 i4 j4  -> I4 J4 i4 j4')
    cUy = Uy.hconjugate('ij|a')
    Yp = E('s i4 j4, i4 j4 lk i2 j2 -> s lk j2 i2', cUy, YY)
    Yp = E('s lk j2 i2, i2 j2 t -> lk t s', Yp, Uy)
    Yp = E('nak, lk t s -> n la t s', Q1, Yp)
    Yp = E('nbl, n la t s -> n ba t s', Q2, Yp)
    T = E(' t1 t3 kl m, n lk t2 t4 -> t1 t2 t3 t4 mn ', Xp, Yp)
    err = None
    if error_test:
        Z1 = E('IJIJmn,KLKLmn->mn', T1o, T2o)
        Z2 = E('IJIJmn->mn', T)
        err = (Z1 - Z2).norm / Z1.norm
    T, Tnorm = _normalised(T)
    return (T, Tnorm, err) if error_test else (T, Tnorm)


# ------------------------------------------------------------------------------------------------
#  initial-tensor pipeline: the compression stages of tensor_preparation
#  (reference gauge2d.py:22-66) on the GPU ops.  The symbolic construction of A and B (get_ABtensors: sympy Berezin
#  integrals, reference gauge2d.py:139-257) stays with the reference; its output arrays go in as gtn.dense.
# ------------------------------------------------------------------------------------------------
def _env_isometry(Qp, hp, Qm, hm, gram, part, cutoff):
    """Isometry onto the dominant subspace of a leg: eigen-decompositions of the two environment Gram matrices
    Mp = Qp^dagger Qp and Mm = Qm Qm^dagger (the leg as the last / first group of Qp / Qm); the one with the smaller
    bond wins (reference gauge2d.py:1225-1243, :1313-1338, :1533-1541).  Returns (U, U^dagger)."""
    E = gtn.einsum
    Mp = E(gram, Qp.hconjugate(hp), Qp)
    Mm = E(gram, Qm, Qm.hconjugate(hm))
    Up, Lp, cUp = Mp.eig(part, cutoff)
    Um, Lm, cUm = Mm.eig(part, cutoff)
    return (Up, cUp) if Lp.shape[0] < Lm.shape[0] else (Um, cUm)


def fcompress_B(B, cutoff=64, mute=True):
    """First compression of the link tensor B[I,J,K,L,i,j,k,l] (fermionic legs only; reference gauge2d.py:1198-1291):
    one isometry per direction, applied to the forward leg and, conjugated, to the backward leg."""
    E = gtn.einsum
    g7 = 'I abcdefg,abcdefg J -> IJ'
    U1, _ = _env_isometry(E('IJKLijkl -> JKLijkl I', B), 'abcdefg|x', E('IJKLijkl -> K IJLijkl', B), 'x|abcdefg',
                          g7, 'I|J', cutoff)
    B = E('IA,IJKLijkl->AJKLijkl', U1, B)
    B = E('CK,AJKLijkl->AJCLijkl', U1.hconjugate('I|J'), B)
    U2, cU2 = _env_isometry(E('IJKLijkl -> IKLijkl J', B), 'abcdefg|x', E('IJKLijkl -> L IJKijkl', B), 'x|abcdefg',
                            g7, 'I|J', cutoff)
    B = E('JB,AJCLijkl->ABCLijkl', U2, B)
    B = E('DL,ABCLijkl->ABCDijkl', cU2, B)
    return B


def _boson_delta(n):
    import numpy as _np
    return gtn.dense(_np.eye(n), statistics=(0, 0))


def compress_B(B, cutoff=64, mute=True):
    """Second compression of B: every fermionic leg is joined with its bosonic partner (copied through a delta) and
    truncated (reference gauge2d.py:1293-1467).  Returns (B[A,B,C,D], [U1, U2, U3, U4])."""
    E = gtn.einsum
    d = _boson_delta(B.shape[4])
    g9 = 'Ii abcdefg,abcdefg Jj -> IiJj'
    # per direction: (delta leg, Qp string, Qm string): the reference's four hand-unrolled blocks
    dirs = [('k', 'JKLjklm Ii', 'Kk IJLijlm'), ('l', 'IKLiklm Jj', 'Ll IJKijkm'),
            ('i', 'JKLjklm Ii', 'Kk IJLijlm'), ('j', 'IKLiklm Jj', 'Ll IJKijkm')]
    Us = []
    for leg, sp, sm in dirs:
        Qp = E('IJKLijkl,%sm -> %s' % (leg, sp), B, d)
        Qm = E('IJKLijkl,%sm -> %s' % (leg, sm), B, d)
        U, _ = _env_isometry(Qp, 'JKLjklm|Ii', Qm, 'Kk|IJLijlm', g9, 'Ii|Jj', cutoff)
        Us.append(U)
        del Qp, Qm
    U1, U2, U3, U4 = Us
    Bf = E('IJKLijkl,IiA->AJKLjkl', B, U1)
    Bf = E('AJKLjkl,JjB->ABKLkl', Bf, U2)
    Bf = E('ABKLkl,CKk->ABCLl', Bf, U3.hconjugate('ij|k'))
    Bf = E('ABCLl,DLl->ABCD', Bf, U4.hconjugate('ij|k'))
    return Bf, Us


def compress_A(A, Upack, mute=True):
    """Site tensor A dressed with the isometries of compress_B (reference gauge2d.py:1469-1501)."""
    E = gtn.einsum
    U1, U2, U3, U4 = Upack
    d = _boson_delta(A.shape[0])
    Ix = E('KXj,XjI->KIj', U1.hconjugate('ij|k'), U3)
    Iy = E('LYj,YjJ->LJj', U2.hconjugate('ij|k'), U4)
    return E('ijkl,KIl,LJk,km,ln->IJKLijklmn', A, Ix, Iy, d, d)


def compress_T(T, cutoff=64, mute=True):
    """Final compression of T[I,J,K,L,i,j,k,l,m,n] along x, then y (reference gauge2d.py:1503-1585)."""
    E = gtn.einsum
    U, _ = _env_isometry(E('IJKLijklmn -> JKLjklmn Ii', T), 'abcdefgh|xy', E('IJKLijklmn -> Kk IJLijlmn', T),
                         'xy|abcdefgh', 'Ii abcdefgh, abcdefgh Jj -> IiJj', 'Ii|Jj', cutoff)
    Tf = E('IJKLijklmn,IiA,CKk->ACJLjlmn', T, U, U.hconjugate('ij|k'))
    U, _ = _env_isometry(E('IKJLjlmn -> IKLlmn Jj', Tf), 'abcdef|xy', E('IKJLjlmn -> Ll IKJjmn', Tf), 'xy|abcdef',
                         'Ii abcdef, abcdef Jj -> IiJj', 'Ii|Jj', cutoff)
    return E('ACJLjlmn,JjB,DLl->ABCDmn', Tf, U, U.hconjugate('ij|k'))


def tensor_from_AB(A, B, cutoff=64):
    """tensor_preparation (reference gauge2d.py:22-66) from the A / B tensors on: B compression (1), (2), A compression,
    T formation, T compression.  Returns (T, trace_error) with trace_error = |1 - z4 / z1| of the reference (:32, :57-59)."""
    import numpy as _np
    E = gtn.einsum
    nphi = A.shape[0]
    z1 = E("IJIJijij,jiji", B, A)
    B = fcompress_B(B, cutoff)
    B, Us = compress_B(B, cutoff)
    A = compress_A(A, Us)
    T = E('IJXYijklmn,XYKL->IJKLijklmn', A, B)
    T = compress_T(T, cutoff)
    z4 = E("IJIJij,ij", T, gtn.dense(_np.full((nphi, nphi), 1.0), statistics=(0, 0)))
    return T, abs(1 - z4 / z1)


def logZhotrg3dz(T1, T2, boundary_conditions="periodic"):
    """reference gauge2d.py:1617-1641"""
    if boundary_conditions == "anti-periodic" and not isinstance(T1, gtn.block):
        T1, T2 = _flip_leg_parity(T1, 1), _flip_leg_parity(T2, 1)
    return np.log(gtn.einsum('IJIJmn,KLKLmn', T1, T2))


def coarse_grain(T, cgsteps=5, dcut=32, method="atrg", boundary_conditions="anti-periodic", error_test=False,
                 log=None, checkpoint_dir=None, resume=False, checkpoint_every=1, logNorm0=0.0):
    """The 2D loop of example.py:156-196 (after zcap): returns (T, records) with one record per step
    (volume, F, Tnorm, err, shape).
    log: path of a JSON-lines run log (checkpoint.RunLog), one line per step.
    checkpoint_dir: the tensor and the accumulated log-norm are saved after every `checkpoint_every` steps;
    with resume=True the loop continues from the newest checkpoint found there with the same tensor bits (the
    continuation agrees with an uninterrupted run to rounding; the reference has neither).  logNorm0: log-norm accumulated before the 2D loop (flavour coarse-graining,
    example.py:144-150)."""
    import time as _time
    from . import checkpoint as ck
    logNorm = float(logNorm0)
    records = []
    runlog = ck.RunLog(log, truncate=not resume) if log else None
    first = 0
    cgxfirst = T.shape[0] > T.shape[1]
    if checkpoint_dir:
        os.makedirs(checkpoint_dir, exist_ok=True)
        step, path = ck.latest_step(checkpoint_dir) if resume else (None, None)
        if step is not None:
            T, meta = ck.load_tensor(path)
            have = (meta["dcut"], meta["method"], meta["boundary_conditions"], bool(meta.get("error_test", error_test)),
                    float(meta.get("logNorm0", logNorm0)))
            if have != (dcut, method, boundary_conditions, bool(error_test), float(logNorm0)):
                gtn.error("Error[coarse_grain]: the checkpoint was written with different run parameters.")
            logNorm, first, cgxfirst, records = meta["logNorm"], step, meta["cgxfirst"], meta["records"]
            for r in records:
                r["F"], r["shape"] = complex(*r["F"]), tuple(r["shape"])
            if runlog:
                runlog.truncate_to(len(records))        # steps logged after the checkpoint are run (and logged) again
    if first == 0:
        F = logZ(T, boundary_conditions) + logNorm
        records.append(dict(vol=1, F=complex(F), Tnorm=None, err=None, shape=tuple(getattr(T, "effective_shape", T.shape)[:2])))
        if runlog:
            runlog.write(process="_ini", **records[-1])
    for i in range(first, cgsteps):
        t0 = _time.time()
        if method == "trg":
            process = "_trg"
            out = trg(T, dcut, iternum=i, error_test=error_test)
        else:
            use_x = (i % 2 == 0) == cgxfirst
            fn, process = (atrg2dx, "_atrgx") if use_x else (atrg2dy, "_atrgy")
            out = fn(T, T, dcut, iternum=i, error_test=error_test)
        T, Tnorm = out[0], out[1]
        vol = 2 ** (i + 1)
        logNorm = 2 * logNorm + math.log(Tnorm)
        F = (logZ(T, boundary_conditions) + logNorm) / vol
        records.append(dict(vol=vol, F=complex(F), Tnorm=float(Tnorm), err=(float(out[2]) if error_test else None),
                            shape=tuple(getattr(T, "effective_shape", T.shape)[:2])))
        if runlog:
            runlog.write(process=process, seconds=_time.time() - t0, **records[-1])
        if checkpoint_dir and ((i + 1) % checkpoint_every == 0 or i + 1 == cgsteps):
            meta_records = [dict(r, F=[r["F"].real, r["F"].imag], shape=list(r["shape"])) for r in records]
            ck.save_tensor(ck.step_path(checkpoint_dir, i + 1), T, logNorm=logNorm, dcut=dcut, method=method,
                           boundary_conditions=boundary_conditions, cgxfirst=bool(cgxfirst), records=meta_records,
                           error_test=bool(error_test), logNorm0=float(logNorm0))
    return T, records
