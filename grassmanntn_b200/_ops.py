"""Grassmann einsum / decomposition / conjugation on parity-blocked device tensors (BT).

Host logic only decides, per parity block, where the block goes, its scalar sign and which legs
carry a sigma vector; the data movement and arithmetic are three kernels:
  pack   (gtn_sign_permute)  operand blocks  -> packed [rows x cols] matrices, rows/cols ordered
                              by (total parity, parity pattern, offset)
  gemm   (gtn_grouped_gemm)   one DMMA GEMM per output parity block, K restricted to the parity
                              sector that can be non-zero for Grassmann-even operands
  svd    (gtn_jacobi_*)       batched one-sided Jacobi on the even/odd sector matrices
References: einsum_ds / einsum_block (reference __init__.py:1633-2306, :2346-2941), svd / eig /
decompose_block (:4033-4308, :4425-4698, :4704-5289), hconjugate(_block) (:5300-5493, :5495-5955).
"""
import math
import os
import sys

import numpy as np
import torch

from . import _planner as P
from . import parallel
from ._engine import (BT, FERMI, GemmPlan, GroupLayout, group_layout, PermutePlan, _cached, _ptr, _row_strides, _stream,
                      batched_svd, truncated_svd_batch, bt_force_standard, bt_switch_format, build_job, dtype_code, gemm, lin_leg,
                      require_cuda, sigma_bits)
from ._cabi import check, count, lib
from . import _engine as _engine_mod

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_CORE_FILES = ("__init__.py", "_ops.py", "_engine.py")
from ._engine import prof_region

COMPACT_PACK = bool(int(os.environ.get("GTN_COMPACT_PACK", "1")))
NUMER_CUTOFF = 1.0e-14      # reference __init__.py:30 (module global numer_cutoff)
SVD_PATH_STATS = {"truncated": 0, "truncated_rejected": 0, "full": 0}
TRUNCATED_SVD = True        # randomized subspace path for truncated decompositions (with certificate)


def _err(msg):
    raise P.GtnValueError(msg)


# ------------------------------------------------------------------------------------------------
#  unwritten permutations (no permuted copy between two ops that both move the data)
# ------------------------------------------------------------------------------------------------
LAZY_PERMUTE = bool(int(os.environ.get("GTN_LAZY_PERMUTE", "1")))
LAZY_STATS = {"created": 0, "fused": 0, "materialised": 0}


class LazyPermute:
    """A tensor that is a signed leg permutation (optionally conjugated) of a stored tensor `src` and has not been
    written: leg p of the tensor is leg perm[p] of src, the value carries (-1)^e with e the GF(2) form
    sum_{a in alpha} p_a + sum_{a in beta} sigma_a + sum_{{a,b} in Q} p_a p_b over the legs of src.
    The reference writes every such intermediate (T1 = einsum('ijkl->jkli', T) at gauge2d.py:1673, the leg order of
    VV / UU at :1727-1732, the conjugates at :1974 / :1985) and the consumer then moves the data once more to build
    its matrix; Grassmann reorderings compose (the exponents add), so a consumer that packs anyway (PackPlan,
    the matricisation of a decomposition) reads src through the composed tables: one pass and one launch less per
    pair.  Any other access (`BT.buf`) writes the tensor first, with the launch plan it was created with."""
    __slots__ = ("src", "perm", "alpha", "beta", "Q", "conj", "sig", "_mat", "__weakref__")

    def __init__(self, src, perm, alpha, beta, Q, conj, mat):
        assert src.pending() is None
        self.src, self.perm = src, tuple(perm)
        self.alpha, self.beta, self.Q = frozenset(alpha), frozenset(beta), frozenset(Q)
        self.conj = bool(conj)
        self.sig = ("lazy", src.key(), self.perm, tuple(sorted(self.alpha)), tuple(sorted(self.beta)),
                    tuple(sorted(tuple(sorted(pr)) for pr in self.Q)), self.conj)
        self._mat = mat                      # (source buffer, scale) -> the tensor's buffer

    def materialise(self, scale=1.0):
        LAZY_STATS["materialised"] += 1
        return self._mat(self.src.buf, scale)


def _attach_lazy(res, lz):
    """make the (buffer-less) BT `res` stand for lz; src remembers it so that handing out writable views of src
    (block.data) can write its dependents first"""
    res._buf, res._lazy = None, lz
    res.__dict__.pop("_key", None)
    _engine_mod._register_dependent(lz.src, res)
    LAZY_STATS["created"] += 1
    return res


def settle(bt):
    """write every unwritten permutation of bt (call before bt's storage is modified in place)"""
    for d in list(bt.__dict__.get("_deps", ())):
        d.buf
    bt.__dict__.pop("_deps", None)


def _through_lazy(bt, labels, assigned):
    """the same pack expressed on the source of an unwritten permutation: (src, labels of src's legs, total sign
    assignment, conj)"""
    lz = bt.pending()
    inv = [0] * len(lz.perm)
    for p_, a in enumerate(lz.perm):
        inv[a] = p_
    src_labels = [labels[inv[a]] for a in range(len(lz.perm))]
    alpha, beta, Q = set(assigned[0]), set(assigned[1]), set(assigned[2])
    for a in lz.alpha:
        alpha ^= {src_labels[a]}
    for a in lz.beta:
        beta ^= {src_labels[a]}
    for pr in lz.Q:
        a, b = tuple(pr)
        x, y = src_labels[a], src_labels[b]
        if x == y:
            alpha ^= {x}                 # both legs on the diagonal of a trace: p * p = p
        else:
            Q ^= {frozenset((x, y))}
    return lz.src, src_labels, (alpha, beta, Q), lz.conj


# ------------------------------------------------------------------------------------------------
#  packing an operand into [batch][rows][cols][t]
# ------------------------------------------------------------------------------------------------
def _label_legs(bt, labels):
    info = {}
    for a, ch in enumerate(labels):
        leg = bt.leg(a)
        if ch in info:
            if info[ch][1:] != leg[1:]:
                _err("Error[einsum]: index '%s' has inconsistent dimensions" % ch)
        else:
            info[ch] = leg
    return info


def _compact_sizes(layR, layC):
    """(elements per batch index, base of the odd sector) of the compact packing of an even operand:
    [A_ee (R_e x C_e)][A_oo (R_o x C_o)] back to back -- the zero quadrants of [[A_ee, 0], [0, A_oo]] are not stored"""
    (_, re_), (_, ro) = layR.sector(0), layR.sector(1)
    (_, ce), (_, co) = layC.sector(0), layC.sector(1)
    return re_ * ce + ro * co, re_ * ce


def _pack_jobs(bt, labels, info, groups, lays, assigned, out_elems_mult=None, compact=False, conj=False):
    """Jobs that copy every live block of `bt` into the packed buffer.
    groups: dict name -> list of labels for 'B' (batch), 'R', 'C', 'T'; lays: GroupLayout per name.
    assigned: (alpha labels, beta labels, Q pairs) evaluated by this operand.
    compact (Grassmann-even operand, no traced legs): the two parity sectors are stored back to back as dense
    R_s x C_s matrices (half the bytes of the full R x C matrix; at D = chi = 256 the difference is 32 GiB)."""
    alpha_l, beta_l, q_pairs = assigned
    tot = {k: lays[k].total for k in lays}
    mult = {"T": 1, "C": tot["T"], "R": tot["T"] * tot["C"], "B": tot["T"] * tot["C"] * tot["R"]}
    if compact:
        csz, codd = _compact_sizes(lays["R"], lays["C"])
    uniq = list(dict.fromkeys(labels))
    where = {}
    for g, lst in groups.items():
        for pos, ch in enumerate(lst):
            where[ch] = (g, pos)
    jobs = []
    used = set()
    for pat in bt.live():
        pis = dict(zip(bt.faxes, pat))
        par, ok = {}, True
        for a, ch in enumerate(labels):
            if a in pis:
                if ch in par and par[ch] != pis[a]:
                    ok = False           # traced pair with different parities: off the diagonal
                    break
                par[ch] = pis[a]
        if not ok:
            continue
        gp = {g: tuple(par[ch] for ch in lst if info[ch][0] in FERMI) for g, lst in groups.items()}
        used.add(tuple(gp[g] for g in ("B", "R", "C", "T")))
        out_base = sum(lays[g].offset[gp[g]] * mult[g] for g in lays)
        if compact:
            sec = sum(gp["R"]) % 2
            assert sum(gp["C"]) % 2 == sec, "compact packing needs a Grassmann-even operand"
            (r0, _), (c0, cl) = lays["R"].sector(sec), lays["C"].sector(sec)
            mult = {"T": 1, "C": 1, "R": cl, "B": csz}
            out_base = (lays["B"].offset[gp["B"]] * csz + (codd if sec else 0)
                        + (lays["R"].offset[gp["R"]] - r0) * cl + (lays["C"].offset[gp["C"]] - c0))
        bshape = bt.block_shape(pat)
        bstr = _row_strides(bshape)
        legs, beta = [], []
        for ch in uniq:
            axes = [a for a, c in enumerate(labels) if c == ch]
            n = bshape[axes[0]]
            g, pos = where[ch]
            ostr = lays[g].strides(gp[g])[pos] * mult[g]
            q = None
            if ch in beta_l:
                q = sigma_bits(par[ch], n)
            legs.append(lin_leg(n, sum(bstr[a] for a in axes), ostr, q=q))
            beta.append(1 if ch in beta_l else 0)
        const = 0
        for x in alpha_l:
            const ^= par[x]
        for pr in q_pairs:
            x, y = tuple(pr)
            const ^= par[x] & par[y]
        jobs.append(build_job(legs, beta=beta, const=const, conj=conj, in_base=bt.off[pat], out_base=out_base))
    return jobs, used


def _alloc(fn, n, dtype, dev):
    """large temporaries: when the caching allocator cannot serve the request from its free pool, return the cached
    blocks to the driver and try once more (a 64 GiB pack at D = chi = 256 next to 30 GiB of cached, unused blocks)"""
    try:
        return fn(n, dtype=dtype, device=dev)
    except torch.OutOfMemoryError:
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        return fn(n, dtype=dtype, device=dev)


class PackPlan:
    """cached launch plan that packs every live block of an operand layout into [B][R][C][T]
    (then sums the T axis away)."""

    def __init__(self, bt, labels, info, groups, assigned, restrict_even=None, compact=False):
        self.lays = lays = {g: group_layout([info[ch] for ch in groups[g]]) for g in ("B", "R", "C", "T")}
        self.compact = bool(compact and restrict_even and COMPACT_PACK and not groups["T"] and bt.is_even())
        self.fused = bt.pending() is not None
        conj = False
        if self.fused:
            # the operand is an unwritten permutation: pack straight from its source through the composed tables
            bt, labels, assigned, conj = _through_lazy(bt, labels, assigned)
            LAZY_STATS["fused"] += 1
        self.src_labels, self.assigned, self.conj = list(labels), assigned, conj
        jobs, used = _pack_jobs(bt, labels, info, groups, lays, assigned, compact=self.compact, conj=conj)
        self.plan = PermutePlan(jobs)
        self.total = lays["B"].total * lays["R"].total * lays["C"].total * lays["T"].total
        if self.compact:
            self.total = lays["B"].total * _compact_sizes(lays["R"], lays["C"])[0]
        need = 1
        for g in ("B", "R", "C", "T"):
            need *= len(lays[g].pats)
        self.even = bt.is_even() if restrict_even is None else restrict_even
        if self.even:
            # consumers of an even operand only touch parity-matched sectors
            readable = sum(1 for pb in lays["B"].pats for pr in lays["R"].pats for pc in lays["C"].pats
                           for pt in lays["T"].pats if (sum(pr) + sum(pc)) % 2 == 0)
            self.full = len(used) >= readable
        else:
            self.full = len(used) >= need
        self.reduce = lays["T"].total > 1 or bool(groups["T"])
        self.rows = lays["B"].total * lays["R"].total * lays["C"].total
        self.tcols = lays["T"].total

    def run(self, bt):
        return self.run_src(bt.pending().src.buf if self.fused else bt.buf)

    def run_src(self, src, scale=1.0):
        dev = src.device
        buf = _alloc(torch.empty if self.full else torch.zeros, max(self.total, 1), src.dtype, dev)
        self.plan.run(src, buf, scale)
        if self.reduce:
            red = torch.empty(max(self.rows, 1), dtype=src.dtype, device=dev)
            check(lib.gtn_rowsum(_ptr(buf), _ptr(red), self.rows, self.tcols, dtype_code(src.dtype), _stream()),
                  "gtn_rowsum")
            count()
            buf = red
        return buf


def _pack(bt, labels, info, groups, assigned, restrict_even=None):
    """returns (buffer [B*R*C] after the t-sum, layouts, even flag)."""
    pp = PackPlan(bt, labels, info, groups, assigned, restrict_even)
    return pp.run(bt), pp.lays, pp.even


def _out_bt(info, out_labels, dtype, lays_order, pats_offsets):
    stats = [info[ch][0] for ch in out_labels]
    e = [info[ch][1] for ch in out_labels]
    o = [info[ch][2] for ch in out_labels]
    bt = BT(stats, e, o, dtype)
    bt.off = dict(pats_offsets)
    return bt


# ------------------------------------------------------------------------------------------------
#  einsum
# ------------------------------------------------------------------------------------------------
_einsum_cache = {}


def einsum_bt(subscripts, ops, ignore_anticommutation=False):
    ops = [bt_force_standard(o) for o in ops]
    dt = torch.complex128 if any(o.dtype == torch.complex128 for o in ops) else torch.float64
    ops = [o if o.dtype == dt else _cast(o, dt) for o in ops]
    key = (subscripts, bool(ignore_anticommutation), tuple(o.key() for o in ops))
    ex = _einsum_cache.get(key)
    if ex is None:
        if len(_einsum_cache) > 2048:
            _einsum_cache.clear()
            _engine_mod.drop_graphs()        # recorded graphs point into the launch tables of the evicted executors
        ex = _build_einsum(subscripts, ops, ignore_anticommutation)
        _einsum_cache[key] = ex
    return ex(ops)


def _build_einsum(subscripts, ops, ignore_anticommutation):
    inputs, output = P.parse_subscripts(subscripts)
    if len(inputs) != len(ops):
        _err("Error[einsum]: the number of subscripts does not match the number of operands")
    for sub, o in zip(inputs, ops):
        if len(sub) != o.ndim:
            _err("Error[einsum]: subscript '%s' does not match a tensor with %d legs" % (sub, o.ndim))
    if len(ops) > 2:
        return lambda ops_: _einsum_fold(inputs, output, ops_, ignore_anticommutation)
    prog, first_stat, contracted = P.einsum_sign_program(inputs, output, [o.stats for o in ops],
                                                         ignore_anticommutation)
    if len(ops) == 1:
        return _plan_single(inputs[0], output, ops[0], prog)
    return _plan_pair(inputs, output, ops, prog)


def _cast(bt, dt):
    r = BT(bt.stats, bt.e, bt.o, dt, bt.fmt)
    r.off, r.zero = dict(bt.off), set(bt.zero)
    r.buf = bt.buf.to(dt)
    return r


def _einsum_fold(inputs, output, ops, ignore):
    """N >= 3 operands: contract left to right (Grassmann contraction is associative as long as
    the operands stay in order); every pairwise step is a full Grassmann einsum."""
    cur_sub, cur = inputs[0], ops[0]
    for k in range(1, len(ops)):
        rest = "".join(inputs[k + 1:]) + (output or "")
        both = cur_sub + inputs[k]
        keep = []
        for ch in dict.fromkeys(both):
            cnt = both.count(ch)
            st = (cur.stats[cur_sub.index(ch)] if ch in cur_sub else ops[k].stats[inputs[k].index(ch)])
            if ch in rest:
                keep.append(ch)
            elif st == 0 and cnt == 1:
                pass           # lone bosonic index summed away
            elif cnt == 1:
                keep.append(ch)
        if k == len(ops) - 1:
            sub = cur_sub + "," + inputs[k] + ("->" + output if output is not None else "")
        else:
            sub = cur_sub + "," + inputs[k] + "->" + "".join(keep)
        cur = einsum_bt(sub, [cur, ops[k]], ignore)
        cur_sub = "".join(keep)
    return cur


def _scalar(buf):
    v = buf[:1].cpu().numpy()[0]
    return v


def _bt_from_layout(info, labels, dtype, lay, even, buf):
    res = _out_bt(info, labels, dtype, None, {})
    for p in lay.pats:
        if even and sum(p) % 2 == 1:
            continue
        if lay.size[p] > 0:
            res.off[p] = lay.offset[p]
    # the layout is parity-sorted: for an even tensor the (never written) odd part is the tail
    if buf is not None:
        res.buf = buf[: max(lay.even_total, 1)] if even else buf
    return res


def _lazy_result(pp, info, out, x):
    """the result of a pure permutation whose pack plan is pp, unwritten (LazyPermute); a permutation of an unwritten
    permutation refers to the stored tensor underneath (chains collapse)"""
    lay, even = pp.lays["R"], pp.even
    res = _bt_from_layout(info, out, x.dtype, lay, even, None)
    root = x.pending().src if pp.fused else x
    ax = {ch: a for a, ch in enumerate(pp.src_labels)}
    al, be, Q = pp.assigned

    def mat(src_buf, scale):
        buf = pp.run_src(src_buf, scale)
        return buf[: max(lay.even_total, 1)] if even else buf
    lz = LazyPermute(root, [ax[ch] for ch in out], {ax[c] for c in al}, {ax[c] for c in be},
                     {frozenset(ax[c] for c in pr) for pr in Q}, pp.conj, mat)
    return _attach_lazy(res, lz)


def _plan_single(labels, output, bt, prog):
    info = _label_legs(bt, labels)
    out = output or ""
    t_labels = [ch for ch in dict.fromkeys(labels) if ch not in out]
    for ch in dict.fromkeys(labels):
        if labels.count(ch) == 2 and ch in out:
            raise NotImplementedError("einsum: a repeated index that is kept in the output is not supported")
    groups = {"B": [], "R": list(out), "C": [], "T": t_labels}
    pp = PackPlan(bt, labels, info, groups, (prog.alpha, prog.beta, prog.Q))
    pure = (output is not None and not t_labels and len(set(labels)) == len(labels)
            and len(out) == len(labels) and len(labels) > 0)

    def run(ops):
        if pure and LAZY_PERMUTE:
            return _lazy_result(pp, info, out, ops[0])
        buf = pp.run(ops[0])
        if output is None:
            return _scalar(buf)
        return _bt_from_layout(info, out, ops[0].dtype, pp.lays["R"], pp.even, buf)
    return run


def _plan_pair(inputs, output, ops, prog):
    la, lb = inputs
    A, B = ops
    infoA, infoB = _label_legs(A, la), _label_legs(B, lb)
    info = dict(infoA)
    for ch, leg in infoB.items():
        if ch in info and info[ch][1:] != leg[1:]:
            _err("Error[einsum]: index '%s' has inconsistent dimensions" % ch)
        info.setdefault(ch, leg)
    out = output or ""
    cls = {}
    for ch in dict.fromkeys(la + lb):
        cA, cB, co = la.count(ch), lb.count(ch), out.count(ch)
        if info[ch][0] == 0 and (cA > 1 or cB > 1) and (cA and cB or co):
            # a bosonic index repeated inside one operand addresses that operand's diagonal (the pack tables add the
            # strides of its occurrences), e.g. 'IJIJijij,jiji' of tensor_preparation (reference gauge2d.py:32)
            cA, cB = min(cA, 1), min(cB, 1)
        if cA == 2 and cB == 0 and co == 0:
            cls[ch] = "tA"
        elif cB == 2 and cA == 0 and co == 0:
            cls[ch] = "tB"
        elif cA == 1 and cB == 1 and co == 0:
            cls[ch] = "K"
        elif cA == 1 and cB == 1 and co == 1:
            if info[ch][0] != 0:
                _err("Error[einsum]: Inconsistent index statistics.")
            cls[ch] = "batch"
        elif cA == 1 and cB == 0:
            cls[ch] = "M" if co == 1 else "tA"
        elif cB == 1 and cA == 0:
            cls[ch] = "N" if co == 1 else "tB"
        else:
            raise NotImplementedError("einsum: unsupported index multiplicity for '%s'" % ch)
    M = [ch for ch in out if cls[ch] == "M"]
    N = [ch for ch in out if cls[ch] == "N"]
    batch = [ch for ch in out if cls[ch] == "batch"]
    K = [ch for ch in dict.fromkeys(la) if cls[ch] == "K"]
    tA = [ch for ch in dict.fromkeys(la) if cls[ch] == "tA"]
    tB = [ch for ch in dict.fromkeys(lb) if cls[ch] == "tB"]
    # left operand = the one whose free legs come first in the output
    swap = False
    if M and N and out.index(N[0]) < out.index(M[0]):
        swap = True
    if swap:
        L, R = (B, lb, infoB, N, tB), (A, la, infoA, M, tA)
    else:
        L, R = (A, la, infoA, M, tA), (B, lb, infoB, N, tB)
    Lset, Rset = set(L[1]), set(R[1])
    asg = {"L": (set(), set(), set()), "R": (set(), set(), set())}
    cross = set()
    for x in prog.alpha:
        asg["L" if x in Lset else "R"][0].add(x)
    for x in prog.beta:
        asg["L" if x in Lset else "R"][1].add(x)
    for pr in prog.Q:
        x, y = tuple(pr)
        if x in Lset and y in Lset:
            asg["L"][2].add(pr)
        elif x in Rset and y in Rset:
            asg["R"][2].add(pr)
        else:
            cross.add(pr)
    gL = {"B": batch, "R": L[3], "C": K, "T": L[4]}
    gR = {"B": batch, "R": K, "C": R[3], "T": R[4]}
    even = L[0].is_even() and R[0].is_even()
    ppL = PackPlan(L[0], L[1], info, gL, asg["L"], restrict_even=even, compact=True)
    ppR = PackPlan(R[0], R[1], info, gR, asg["R"], restrict_even=even, compact=True)
    layL, layR = ppL.lays, ppR.lays
    layM, layK, layN, layB = layL["R"], layL["C"], layR["C"], layL["B"]
    Mtot, Ktot, Ntot, Btot = layM.total, layK.total, layN.total, layB.total
    tmp_labels = batch + L[3] + R[3]
    res = _out_bt(info, tmp_labels, A.dtype, None, {})
    # output blocks: pattern over (M fermions, N fermions) in tmp order
    groups, acc = [], 0
    Mferm = [ch for ch in L[3] if info[ch][0] in FERMI]
    Nferm = [ch for ch in R[3] if info[ch][0] in FERMI]
    for pM in layM.pats:
        for pN in layN.pats:
            m, n = layM.size[pM], layN.size[pN]
            if even and (sum(pM) + sum(pN)) % 2 == 1:
                continue
            if m == 0 or n == 0:
                continue
            if even:
                k0, kl = layK.sector(sum(pM) % 2)
            else:
                k0, kl = 0, Ktot
            par = dict(zip(Mferm, pM))
            par.update(zip(Nferm, pN))
            e = 0
            for pr in cross:
                x, y = tuple(pr)
                e ^= par[x] & par[y]
            pat = tuple(pM) + tuple(pN)
            res.off[pat] = acc
            g_ = dict(a_off=layM.offset[pM] * Ktot + k0, b_off=k0 * Ntot + layN.offset[pN], c_off=acc,
                      lda=Ktot, ldb=Ntot, ldc=n, m=m, n=n, k=kl, batch=Btot,
                      bsa=Mtot * Ktot, bsb=Ktot * Ntot, bsc=m * n, alpha=-1.0 if e else 1.0)
            sec = sum(pM) % 2
            if ppL.compact:          # [A_ee][A_oo] back to back: sector-local row offset, lda = K of the sector
                szL, oddL = _compact_sizes(layM, layK)
                g_.update(a_off=(oddL if sec else 0) + (layM.offset[pM] - layM.sector(sec)[0]) * kl, lda=kl, bsa=szL)
            if ppR.compact:
                szR, oddR = _compact_sizes(layK, layN)
                n0, nl = layN.sector(sec)
                g_.update(b_off=(oddR if sec else 0) + (layN.offset[pN] - n0), ldb=nl, bsb=szR)
            groups.append(g_)
            acc += Btot * m * n
    plan = GemmPlan(groups, A.dtype)
    shard_cache = {}
    res_off = dict(res.off)
    res_meta = (res.stats, res.e, res.o)
    tmp = "".join(tmp_labels)
    fin = None
    if output is not None and tmp != out:
        # bring the legs into the requested order (all signs are already applied)
        res.buf = None
        fin = _plan_permute(res, tmp, out)

    def run(ops_):
        A_, B_ = ops_
        Lop, Rop = (B_, A_) if swap else (A_, B_)
        bufL, bufR = ppL.run(Lop), ppR.run(Rop)
        r = BT(res_meta[0], res_meta[1], res_meta[2], A_.dtype)
        r.off = dict(res_off)
        dev_ = bufL.device
        r.buf = _alloc(torch.empty, max(acc, 1), A_.dtype, dev_)
        if Ktot == 0:
            r.buf.zero_()
        elif Mtot == 1 and Ntot == 1 and Btot == 1 and len(groups) == 1:
            # full contraction to a scalar: 1 x K x 1 is a dot product, not a GEMM
            g0 = groups[0]
            nparts = 1184
            part = torch.empty(2 * nparts, dtype=torch.float64, device=dev_)
            with prof_region("dot", 2, 2 * g0["k"] * bufL.element_size()):
                check(lib.gtn_dot(_ptr(bufL[g0["a_off"]:]), _ptr(bufR[g0["b_off"]:]), g0["k"], dtype_code(A_.dtype),
                                  _ptr(r.buf), _ptr(part), nparts, _stream()), "gtn_dot")
            if g0["alpha"] < 0:
                r.buf.neg_()
        elif (parallel.active() and parallel._state["gemm"] and Btot == 1
              and plan.flops >= parallel._state["min_flops"]):
            # output-tile sharding: this rank computes its row range of every output block, then the
            # blocks are completed with one in-place all-gather each (grassmanntn_b200/parallel.py)
            key = (parallel.rank(), parallel.world())
            if key not in shard_cache:
                sg, pieces = parallel.shard_groups(groups, key[0], key[1])
                shard_cache[key] = (GemmPlan([g for g in sg if g["m"] > 0], A_.dtype, config=0), pieces)
            splan, pieces = shard_cache[key]
            # fused: the GEMM epilogue stores the tiles into every rank's output over NVLink peer memory;
            # otherwise GEMM, then one in-place NCCL all-gather per block
            if not (parallel.fused() and all(g.get("beta", 0.0) == 0.0 for g in groups)
                    and parallel.gemm_allgather_fused(splan, bufL, bufR, r.buf)):
                splan.run(bufL, bufR, r.buf)
                parallel.gather_blocks(r.buf, pieces)
        else:
            plan.run(bufL, bufR, r.buf)
        if output is None:
            return _scalar(r.buf)
        return r if fin is None else fin(r)
    return run


def _plan_permute(bt, labels, out_labels):
    info = _label_legs(bt, labels)
    groups = {"B": [], "R": list(out_labels), "C": [], "T": []}
    pp = PackPlan(bt, labels, info, groups, (set(), set(), set()))

    def run(x):
        if LAZY_PERMUTE and len(labels) > 0:
            return _lazy_result(pp, info, out_labels, x)
        return _bt_from_layout(info, out_labels, x.dtype, pp.lays["R"], pp.even, pp.run(x))
    return run



def permute_bt(bt, labels, out_labels):
    """sign-free leg permutation of a BT (one launch)."""
    return _plan_permute(bt, labels, out_labels)(bt)


# ------------------------------------------------------------------------------------------------
#  decompositions
# ------------------------------------------------------------------------------------------------
def _rank_rule(s, cutoff):
    """reference SortedSVD (__init__.py:3935-3944)"""
    if len(s) == 0:
        return 0
    nnz = int(np.sum(np.abs(s / (abs(s[0]) + NUMER_CUTOFF)) > NUMER_CUTOFF))
    if cutoff is not None and cutoff < nnz:
        nnz = cutoff
    return nnz


def _join_sign_terms(legs_idx, stats, fermionic):
    """sign program of joining legs into one leg of statistic -1 in matrix format:
    (-1)^p on the +1 members (get_grouping_sign_factors, __init__.py:3250-3275; join_legs_block
    :3574-3580) times sigma of the joined index = prod sigma_a * (-1)^{sum_{a<b} p_a p_b}
    (switch_format on the joined leg, :3036; join_index's sgn_sigma_perm, :3357-3366)."""
    f = [a for a in legs_idx if fermionic[a]]
    alpha = {a for a in f if stats[a] == 1}
    beta = set(f)
    Q = {frozenset((a, b)) for i, a in enumerate(f) for b in f[i + 1:]}
    return alpha, beta, Q


def _block_jobs(bt, live, alpha, beta, Q, place, through=False):
    """jobs over blocks with leg-level sign terms; place(pat, bshape) -> (out_base, out_strides) or None.
    through: when bt is an unwritten permutation (LazyPermute) the jobs read its SOURCE through the composed sign
    program -- the caller then runs the plan on _src_buf(bt) and takes its live blocks from _live_of(bt)."""
    lz = bt.pending() if through else None
    if lz is not None:
        return _block_jobs_through(bt, lz, live, alpha, beta, Q, place)
    jobs = []
    for pat in live:
        pis = dict(zip(bt.faxes, pat))
        bshape = bt.block_shape(pat)
        pl = place(pat, bshape)
        if pl is None:
            continue
        out_base, ostr, conj = pl
        bstr = _row_strides(bshape)
        legs, bflags = [], []
        for a in range(bt.ndim):
            q = sigma_bits(pis[a], bshape[a]) if a in beta else None
            legs.append(lin_leg(bshape[a], bstr[a], ostr[a], q=q))
            bflags.append(1 if a in beta else 0)
        const = 0
        for a in alpha:
            const ^= pis[a]
        for pr in Q:
            x, y = tuple(pr)
            const ^= pis[x] & pis[y]
        jobs.append(build_job(legs, beta=bflags, const=const, conj=conj, in_base=bt.off[pat], out_base=out_base))
    return jobs


def _block_jobs_through(bt, lz, live, alpha, beta, Q, place):
    """_block_jobs of the tensor bt = (signed permutation lz of lz.src), reading lz.src: leg p of bt is leg lz.perm[p]
    of the source; the sign exponents add (alpha / beta / Q are over bt's legs, lz's over the source's)"""
    src, perm = lz.src, lz.perm
    n = bt.ndim
    A_ = set(perm[a] for a in alpha) ^ set(lz.alpha)
    B_ = set(perm[a] for a in beta) ^ set(lz.beta)
    Q_ = {frozenset((perm[x], perm[y])) for x, y in (tuple(pr) for pr in Q)} ^ set(lz.Q)
    live = set(live)
    jobs = []
    LAZY_STATS["fused"] += 1
    for spat in src.live():
        spis = dict(zip(src.faxes, spat))
        pat = tuple(spis[perm[p]] for p in bt.faxes)
        if pat not in live:
            continue
        bshape = bt.block_shape(pat)
        pl = place(pat, bshape)
        if pl is None:
            continue
        out_base, ostr, conj = pl
        sshape = src.block_shape(spat)
        sstr = _row_strides(sshape)
        legs, bflags = [None] * n, [0] * n
        for p in range(n):
            a = perm[p]
            q = sigma_bits(spis[a], sshape[a]) if a in B_ else None
            legs[a] = lin_leg(sshape[a], sstr[a], ostr[p], q=q)
            bflags[a] = 1 if a in B_ else 0
        const = 0
        for a in A_:
            const ^= spis[a]
        for pr in Q_:
            x, y = tuple(pr)
            const ^= spis[x] & spis[y]
        jobs.append(build_job(legs, beta=bflags, const=const, conj=bool(conj) != lz.conj, in_base=src.off[spat],
                              out_base=out_base))
    return jobs


def _src_buf(bt):
    """the buffer a plan built by _block_jobs(..., through=True) on bt reads"""
    lz = bt.pending()
    return bt.buf if lz is None else lz.src.buf


def _live_of(bt):
    """live blocks of bt; for an unwritten permutation those its source actually holds"""
    lz = bt.pending()
    if lz is None:
        return bt.live()
    have = set()
    for spat in lz.src.live():
        spis = dict(zip(lz.src.faxes, spat))
        have.add(tuple(spis[lz.perm[p]] for p in bt.faxes))
    return [p for p in bt.live() if p in have]


def _decompose_prepare(bt, nl, kind):
    """U, S, V of the (first nl legs | remaining legs) matricisation.
    rule 'dense': reference BlockSVD/BlockEig rank rule (int(cutoff/2) per sector, pad to a power of
    two, __init__.py:3998-4015); rule 'block': decompose_block's (ceil/floor, pad to max, :5076-5103)."""
    this_fmt = bt.fmt
    bt = bt_force_standard(bt)
    n = bt.ndim
    Rl, Cl = list(range(nl)), list(range(nl, n))
    ferm = [s in FERMI for s in bt.stats]
    fR, fC = any(ferm[a] for a in Rl), any(ferm[a] for a in Cl)
    if not bt.is_even():
        # stored odd-parity blocks: the reference decides on their VALUES (mean |x| over the odd entries
        # <= 1e-14, __init__.py:3977-3983); blocks that are numerically zero are flagged and skipped
        odd = [p for p in bt.live() if sum(p) % 2 == 1]
        n_odd = sum(bt.block_size(p) for p in odd)
        if float(bt.sumabs(odd).item()) / max(n_odd, 1) > NUMER_CUTOFF:
            _err("Error[BlockSVD]: This matrix is not constructed from a Grassmann-even tensor.")
        import copy
        bt = copy.copy(bt)
        bt.zero = set(bt.zero) | set(odd)
        bt.__dict__.pop("_key", None)
    layR = group_layout([bt.leg(a) for a in Rl])
    layC = group_layout([bt.leg(a) for a in Cl])
    sectors = [0, 1] if (fR and fC) else [0]
    alpha, beta, Q = (_join_sign_terms(Rl, bt.stats, ferm) if fR else (set(), set(), set()))
    rows = {s: layR.sector(s) for s in (0, 1)}
    cols = {s: layC.sector(s) for s in (0, 1)}
    if not (fR and fC):
        # one side purely bosonic: an even tensor lives entirely in the parity-0 rows / columns
        rows = {0: layR.sector(0)}
        cols = {0: layC.sector(0)}
    live = _live_of(bt)
    dev = require_cuda()
    fpos_R = [a for a in Rl if ferm[a]]
    fpos_C = [a for a in Cl if ferm[a]]

    def pats_of(pat):
        pis = dict(zip(bt.faxes, pat))
        return tuple(pis[a] for a in fpos_R), tuple(pis[a] for a in fpos_C)

    # ---- pack the sector matrices (one launch)
    sizes = {s: (rows[s][1], cols[s][1]) for s in sectors}
    offs, acc = {}, 0
    for s in sectors:
        offs[s] = acc
        acc += sizes[s][0] * sizes[s][1]
    nlive_needed = sum(1 for pr in layR.pats for pc in layC.pats if (sum(pr) + sum(pc)) % 2 == 0)

    def build_pack():
        def place(pat, bshape):
            pr, pc = pats_of(pat)
            s = (sum(pr) % 2) if (fR and fC) else 0
            ld = sizes[s][1]
            base = offs[s] + (layR.offset[pr] - rows[s][0]) * ld + (layC.offset[pc] - cols[s][0])
            sr, sc = layR.strides(pr), layC.strides(pc)
            return base, [x * ld for x in sr] + list(sc), False
        return PermutePlan(_block_jobs(bt, live, alpha, beta, Q, place, through=True))
    plan = _cached(("svdpack", bt.key(), nl), build_pack)
    Mbuf = (torch.empty if len(live) >= nlive_needed else torch.zeros)(max(acc, 1), dtype=bt.dtype, device=dev)
    plan.run(_src_buf(bt), Mbuf)
    mats = [Mbuf[offs[s]: offs[s] + sizes[s][0] * sizes[s][1]].view(sizes[s][0], sizes[s][1]) for s in sectors]
    if kind == "eig":
        for Mx in mats:
            if Mx.shape[0] != Mx.shape[1]:
                _err("Error[SortedEig]: The input matrix is not Hermitian!")
            nrm2 = torch.zeros(2, dtype=torch.float64, device=dev)
            # Hermiticity check ||M - M^H|| / ||M|| <= 1e-14 (reference :4316-4321) on the library's own kernels:
            # [M ; -M^H] (one copy, one conj-transposing sign+permute launch with scale -1), added by gtn_sum_slices
            nn = Mx.shape[0]
            two = torch.empty(2 * nn * nn, dtype=Mx.dtype, device=dev)
            two[: nn * nn].copy_(Mx.reshape(-1))

            def build_ct(nn=nn, cplx=Mx.dtype == torch.complex128):
                legs = [lin_leg(nn, nn, 1), lin_leg(nn, 1, nn)]
                return PermutePlan([build_job(legs, conj=cplx, in_order=[0, 1], out_order=[1, 0])])
            _cached(("eig_herm_ct", nn, str(Mx.dtype)), build_ct).run(two[: nn * nn], two[nn * nn:], scale=-1.0)
            D = torch.empty(nn * nn, dtype=Mx.dtype, device=dev)
            if nn * nn % 2 and Mx.dtype != torch.complex128:        # (gtn_sum_slices needs 16-byte aligned slices)
                D = two[: nn * nn] + two[nn * nn:]
            else:
                check(lib.gtn_sum_slices(_ptr(two), _ptr(D), nn * nn, 2, dtype_code(Mx.dtype), _stream()), "gtn_sum_slices")
                count()
            check(lib.gtn_sumsq(_ptr(D), D.numel(), dtype_code(D.dtype), _ptr(nrm2[0:]), 0, _stream()), "gtn_sumsq")
            check(lib.gtn_sumsq(_ptr(Mx), Mx.numel(), dtype_code(Mx.dtype), _ptr(nrm2[1:]), 0, _stream()), "gtn_sumsq")
            count(2)
            dn, mn = [math.sqrt(v) for v in nrm2.cpu().tolist()]
            if mn >= NUMER_CUTOFF and dn / mn > NUMER_CUTOFF:
                _err("Error[SortedEig]: The input matrix is not Hermitian!")
    return dict(bt=bt, nl=nl, this_fmt=this_fmt, Rl=Rl, Cl=Cl, fR=fR, fC=fC, layR=layR, layC=layC, sectors=sectors,
                rows=rows, cols=cols, alpha=alpha, beta=beta, Q=Q, fpos_R=fpos_R, fpos_C=fpos_C, dev=dev, mats=mats)


def _decompose_finish(ctx, usv, cutoff, kind, rule, spec=False):
    bt, nl, this_fmt = ctx["bt"], ctx["nl"], ctx["this_fmt"]
    Rl, Cl, fR, fC = ctx["Rl"], ctx["Cl"], ctx["fR"], ctx["fC"]
    layR, layC, sectors, rows, cols = ctx["layR"], ctx["layC"], ctx["sectors"], ctx["rows"], ctx["cols"]
    alpha, beta, Q, fpos_R, fpos_C = ctx["alpha"], ctx["beta"], ctx["Q"], ctx["fpos_R"], ctx["fpos_C"]
    dev = ctx["dev"]

    # ---- rank rule
    if len(sectors) == 2:
        if rule == "dense":
            cuts = [None, None] if cutoff is None else [int(cutoff / 2)] * 2
        else:
            cuts = [None, None] if cutoff is None else [int(math.ceil(cutoff / 2)), int(math.floor(cutoff / 2))]
    else:
        cuts = [cutoff]
    # spec: the singular values are still on the device -- assume every sector keeps its full cut (verified later)
    keep = list(cuts) if spec else [_rank_rule(usv[i][1], cuts[i]) for i in range(len(sectors))]
    if len(sectors) == 2:
        d = max(keep)
        if rule == "dense":
            d = int(2 ** math.ceil(np.log2(d))) if d > 0 else 0
        dims_x = (d, d)
    else:
        d = keep[0]
        dims_x = (d, 0)
    bond_stat = (1, -1) if len(sectors) == 2 else (0, 0)

    # signed eigenvalues for eig: L_k = sum_ij s_i V_ij U_jk (reference :4340-4341)
    svals = []
    for i, s in enumerate(sectors):
        U, sv, Vh = usv[i]
        k = keep[i]
        if kind == "eig" and k > 0:
            Uk = U[:, :k].contiguous()
            Vk = Vh[:k, :].contiguous()
            VU = gemm(Vk.view(-1), Uk.view(-1), k, k, U.shape[0]).view(k, k).cpu().numpy()
            lam = np.einsum("i,ik->k", sv[:k].astype(VU.dtype), VU)
            svals.append(lam)
        else:
            svals.append(sv[:k])

    # ---- U tensor: legs R + x
    Ust = tuple(bt.stats[a] for a in Rl) + (bond_stat[0],)
    Ue = tuple(bt.e[a] for a in Rl) + (dims_x[0],)
    Uo = tuple(bt.o[a] for a in Rl) + (dims_x[1],)
    Ubt = BT(Ust, Ue, Uo, bt.dtype)
    Vst = (bond_stat[1],) + tuple(bt.stats[a] for a in Cl)
    Ve = (dims_x[0],) + tuple(bt.e[a] for a in Cl)
    Vo = (dims_x[1],) + tuple(bt.o[a] for a in Cl)
    Vbt = BT(Vst, Ve, Vo, bt.dtype)
    two = len(sectors) == 2
    if two:
        upats = [tuple(pr) + (sum(pr) % 2,) for pr in layR.pats]
        vpats = [(sum(pc) % 2,) + tuple(pc) for pc in layC.pats]
    else:
        upats = [tuple(pr) for pr in layR.pats if sum(pr) % 2 == 0]
        vpats = [tuple(pc) for pc in layC.pats if sum(pc) % 2 == 0]
    Ubt.alloc([p for p in upats if Ubt.block_size(p) > 0], zero=True)
    Vbt.alloc([p for p in vpats if Vbt.block_size(p) > 0], zero=True)
    # U: one launch per sector source buffer (different base pointers)
    for i, s in enumerate(sectors):
        U, sv, Vh = usv[i]
        k = keep[i]
        if k == 0:
            continue
        if kind == "eig":
            Vh = U.conj().transpose(0, 1)
        Uc = U.contiguous()
        ldu = Uc.shape[1]
        Vc = Vh.contiguous()
        ldv = Vc.shape[1]
        # -- U blocks of this sector
        src = BT(tuple(bt.stats[a] for a in Rl) + (0,), tuple(bt.e[a] for a in Rl) + (k,),
                 tuple(bt.o[a] for a in Rl) + (0,), bt.dtype)
        src.buf = Uc.view(-1)
        for pr in layR.pats:
            if (sum(pr) % 2 if two else 0) != s or layR.size[pr] == 0:
                continue
            if not two and sum(pr) % 2 == 1:
                continue
            src.off[pr] = (layR.offset[pr] - rows[s][0]) * ldu

        def build_u(src=src, s=s, k=k, ldu=ldu):
            jobs = []
            for pr in src.off:
                shp = layR.shape[pr]
                pis = dict(zip(fpos_R, pr))
                upat = tuple(pr) + ((s,) if two else ())
                ostr = _row_strides(Ubt.block_shape(upat))
                sr = layR.strides(pr)
                legs, bflags = [], []
                for j, a in enumerate(Rl):
                    q = sigma_bits(pis[a], shp[j]) if a in beta else None
                    legs.append(lin_leg(shp[j], sr[j] * ldu, ostr[j], q=q))
                    bflags.append(1 if a in beta else 0)
                legs.append(lin_leg(k, 1, ostr[len(Rl)]))
                bflags.append(0)
                const = 0
                for a in alpha:
                    const ^= pis[a]
                for prr in Q:
                    x, y = tuple(prr)
                    const ^= pis[x] & pis[y]
                jobs.append(build_job(legs, beta=bflags, const=const, in_base=src.off[pr], out_base=Ubt.off[upat]))
            return PermutePlan(jobs)
        _cached(("svdU", bt.key(), nl, s, k, ldu, dims_x, kind), build_u).run(src.buf, Ubt.buf)

        # -- V blocks: V_std[x, C] = sigma(x) * Vh[x, C]   (x is the conjugated (-1) leg)
        def build_v(s=s, k=k, ldv=ldv):
            jobs = []
            for pc in layC.pats:
                if (sum(pc) % 2 if two else 0) != s or layC.size[pc] == 0:
                    continue
                if not two and sum(pc) % 2 == 1:
                    continue
                shp = layC.shape[pc]
                vpat = ((s,) if two else ()) + tuple(pc)
                ostr = _row_strides(Vbt.block_shape(vpat))
                sc = layC.strides(pc)
                legs = [lin_leg(k, ldv, ostr[0], q=sigma_bits(s, k) if two else None)]
                bflags = [1 if two else 0]
                for j in range(len(Cl)):
                    legs.append(lin_leg(shp[j], sc[j], ostr[1 + j]))
                    bflags.append(0)
                jobs.append(build_job(legs, beta=bflags, in_base=layC.offset[pc] - cols[s][0], out_base=Vbt.off[vpat]))
            return PermutePlan(jobs)
        _cached(("svdV", bt.key(), nl, s, k, ldv, dims_x, kind), build_v).run(Vc.view(-1), Vbt.buf)

    # ---- S tensor (tiny): diag(sigma(x) * s) per sector, built on the host
    Sbt = BT((bond_stat[1], bond_stat[0]), (dims_x[0],) * 2, (dims_x[1],) * 2, bt.dtype)
    spats = [(0, 0), (1, 1)] if two else [()]
    Sbt.alloc([p for p in spats if Sbt.block_size(p) > 0], zero=True)
    if spec:
        # diag(sigma(x) * s) scattered on the device from the device-resident Ritz values
        for i, s in enumerate(sectors):
            k = keep[i]
            p = (s, s) if two else ()
            if k == 0 or p not in Sbt.off:
                continue
            dd = dims_x[s] if two else dims_x[0]

            def build_idx(s=s, k=k, dd=dd, p=p):
                sg = (1 - 2 * sigma_bits(s, k).astype(np.int64)) if two else np.ones(k, dtype=np.int64)
                return (torch.from_numpy(Sbt.off[p] + np.arange(k, dtype=np.int64) * (dd + 1)).to(dev),
                        torch.from_numpy(sg.astype(np.float64)).to(dev))
            idx, sgd = _cached(("specS", two, s, k, dd, Sbt.off[p], str(dev)), build_idx)
            Sbt.buf.index_copy_(0, idx, (usv[i][1][:k] * sgd).to(Sbt.buf.dtype))
    host = np.zeros(max(Sbt.buf.numel(), 1) if not spec else 0, dtype=np.complex128 if bt.dtype == torch.complex128 else np.float64)
    for i, s in enumerate(sectors if not spec else []):
        k = keep[i]
        p = (s, s) if two else ()
        if k == 0 or p not in Sbt.off:
            continue
        dd = dims_x[s] if two else dims_x[0]
        sg = (1 - 2 * sigma_bits(s, k).astype(np.int64)) if two else np.ones(k, dtype=np.int64)
        vals = np.asarray(svals[i]) * sg
        if host.dtype == np.float64:
            vals = np.real(vals)
        idx = Sbt.off[p] + np.arange(k) * (dd + 1)
        host[idx] = vals
    if not spec and Sbt.buf.numel() > 0 and host.size > 0:
        Sbt.buf.copy_(torch.from_numpy(host[: Sbt.buf.numel()]))
    outs = [Ubt, Sbt, Vbt]
    if this_fmt == "matrix":
        outs = [bt_switch_format(x) for x in outs]
    return outs[0], outs[1], outs[2], tuple(keep)




def decompose_many(items, cutoff, kind, rule, speculative=False, resume=None, site=None):
    """items: list of (bt, nl).  All sector matrices of all items go through ONE batched Jacobi
    run (the two SVDs of a TRG step share their sweeps).
    speculative=True returns (outs, pending): when the truncated SVD could replay its steady-state CUDA graph,
    U, S, V are assembled on the device under the assumption 'certificate passes, every sector keeps its full
    cut' WITHOUT synchronising; the caller enqueues the rest of its step and must call pending.verify()
    (False -> repeat the decomposition with resume=pending: the iteration continues from the workspace state
    of the failed run).  pending is None when nothing was speculated."""
    ctxs = [_decompose_prepare(bt, nl, kind) for bt, nl in items]
    mats = [m for c in ctxs for m in c["mats"]]
    # call site (first frame outside the package core): the truncated SVD remembers its converged iteration
    # count per (batch shape, call site) -- the three SVDs of an ATRG step have the same shapes but very
    # different spectra
    f = sys._getframe(1)
    while f is not None and os.path.basename(f.f_code.co_filename) in _CORE_FILES \
            and os.path.dirname(f.f_code.co_filename) == _PKG_DIR:
        f = f.f_back
    _engine_mod.SVD_SITE[0] = (f.f_code.co_filename, f.f_lineno) if f is not None else None
    if site is not None:
        _engine_mod.SVD_SITE[0] = site          # explicit tag (a driver with several decompositions behind one helper)
    if parallel.active():
        # always through the owner/broadcast path in multi-GPU mode: the replicated operands of the
        # sharded contractions must be bit-identical on every rank (SVD gauges are not unique)
        usv = _svd_distributed(mats, ctxs, cutoff, kind, rule)
        outs, k = [], 0
        for c in ctxs:
            n = len(c["mats"])
            outs.append(_decompose_finish(c, usv[k:k + n], cutoff, kind, rule))
            k += n
        return (outs, None) if speculative else outs
    usv, pending = _svd_local(mats, ctxs, cutoff, kind, rule, speculative, resume)
    outs, k = [], 0
    for c in ctxs:
        n = len(c["mats"])
        outs.append(_decompose_finish(c, usv[k:k + n], cutoff, kind, rule, spec=pending is not None))
        k += n
    return (outs, pending) if speculative else outs


def _sector_cuts(ctxs, cutoff, rule):
    ks = []
    for c in ctxs:
        if len(c["sectors"]) == 2:
            ks += ([int(cutoff / 2)] * 2 if rule == "dense" else
                   [int(math.ceil(cutoff / 2)), int(math.floor(cutoff / 2))])
        else:
            ks += [cutoff]
    return ks


def _svd_distributed(mats, ctxs, cutoff, kind, rule):
    """sector problems are independent: problem i is solved by rank i % W, the isometries are
    broadcast from their owners"""
    w, r = parallel.world(), parallel.rank()
    ks = _sector_cuts(ctxs, cutoff, rule) if cutoff is not None else [None] * len(mats)
    mine = [i for i in range(len(mats)) if i % w == r]
    res = {}
    if mine:
        sub, _ = _svd_core([mats[i] for i in mine], [ks[i] for i in mine], cutoff, kind)
        res = {i: sub[j] for j, i in enumerate(mine)}
    return parallel.broadcast_usv(res, len(mats), mats[0].device, mats[0].dtype)


def _svd_local(mats, ctxs, cutoff, kind, rule, speculative=False, resume=None):
    ks = _sector_cuts(ctxs, cutoff, rule) if cutoff is not None else [None] * len(mats)
    return _svd_core(mats, ks, cutoff, kind, speculative, resume)


def _svd_core(mats, ks, cutoff, kind, speculative=False, resume=None):
    """returns (usv, pending): pending is an _engine.SpeculativeSVD when the truncated path was replayed
    without reading its certificate back (usv[b][1] is then a DEVICE vector of Ritz values)"""
    usv = None
    if cutoff is not None and kind in ("svd", "eig") and TRUNCATED_SVD:
        # (eig: the Hermitian sector's SVD U S Vh from the same truncated solver; the signed eigenvalues
        # lam_k = sum_i s_i (Vh U)_ik are formed from its triplets in _decompose_finish, reference :4340-4341)
        from ._engine import TRUNC_LMAX as LM, subspace_rows
        if all(k >= 1 and 3 * k // 2 + 8 <= LM and 4 * min(subspace_rows(k), LM) <= min(m.shape) for k, m in zip(ks, mats)):
            usv = truncated_svd_batch(mats, ks, speculative=speculative, resume=resume)
            if isinstance(usv, _engine_mod.SpeculativeSVD):
                SVD_PATH_STATS["truncated"] += 1
                return usv.outs, usv
            if usv is None and not _engine_mod.CAPTURING_STEP[0] \
                    and min(min(m.shape) for m in mats) >= _engine_mod.ROBUST_MIN_DIM:
                # the spectrum inside the cut spans more than a Gram matrix resolves (s_k < 3e-7 s_0: the whitening
                # dropped directions) or the iteration stalled: before paying for the full SVD of a large sector,
                # repeat the iteration with its panels orthonormalised by the Jacobi kernels
                usv = truncated_svd_batch(mats, ks, robust=True)
                if usv is not None:
                    SVD_PATH_STATS["truncated_robust"] = SVD_PATH_STATS.get("truncated_robust", 0) + 1
                    pk = (tuple(m.shape[0] for m in mats), tuple(m.shape[1] for m in mats), tuple(ks),
                          str(mats[0].dtype), str(mats[0].device))
                    _engine_mod._trunc_robust[(pk, _engine_mod.SVD_SITE[0])] = True
            SVD_PATH_STATS["truncated" if usv is not None else "truncated_rejected"] += 1
    if usv is None:
        if _engine_mod.CAPTURING_STEP[0]:
            raise _engine_mod.NotCapturable("the full Jacobi SVD synchronises with the host")
        usv = batched_svd(mats)
        usv = [_engine_mod.refine_null_band(m, r) for m, r in zip(mats, usv)]
        SVD_PATH_STATS["full"] += 1
    return usv, None


def decompose_bt(bt, nl, cutoff, kind, rule):
    return decompose_many([(bt, nl)], cutoff, kind, rule)[0]


def hconjugate_bt(bt, nl):
    """Hermitian conjugate of the (first nl | rest) matricisation, one launch.
    reference hconjugate (__init__.py:5300-5493) / hconjugate_block (:5495-5955)."""
    this_fmt = bt.fmt
    bt = bt_force_standard(bt)
    n = bt.ndim
    Rl, Cl = list(range(nl)), list(range(nl, n))
    ferm = [s in FERMI for s in bt.stats]
    new_order = Cl + Rl
    flip = lambda s: -s if s in FERMI else s
    res = BT([flip(bt.stats[a]) for a in new_order], [bt.e[a] for a in new_order], [bt.o[a] for a in new_order],
             bt.dtype)
    # join signs of the old row group (stats as they are) ...
    a1, b1, q1 = _join_sign_terms(Rl, bt.stats, ferm)
    # ... and split signs of the new row group = old column group with FLIPPED statistics
    fl = {a: flip(bt.stats[a]) for a in range(n)}
    a2, b2, q2 = _join_sign_terms(Cl, fl, ferm)
    alpha, beta, Q = a1 ^ a2, b1 ^ b2, q1 ^ q2
    live = _live_of(bt)
    npat = lambda pat: tuple(dict(zip(bt.faxes, pat))[a] for a in new_order if ferm[a])
    total = res.plan_blocks([npat(p) for p in live])

    def build():
        def place(pat, bshape):
            op = npat(pat)
            ostr_new = _row_strides(res.block_shape(op))
            ostr = [0] * n
            for newpos, a in enumerate(new_order):
                ostr[a] = ostr_new[newpos]
            return res.off[op], ostr, True
        return PermutePlan(_block_jobs(bt, live, alpha, beta, Q, place, through=True))
    plan = _cached(("hconj", bt.key(), nl), build)

    def mat(src_buf, scale):
        out = _alloc(torch.empty, max(total, 1), src_buf.dtype, src_buf.device)
        plan.run(src_buf, out, scale)
        return out
    res.zero = set()
    if LAZY_PERMUTE and this_fmt != "matrix" and n > 0:
        # the conjugate is a signed, conjugated permutation: left unwritten for the contraction that packs it
        # (hotrg3dz: cA = AA.hconjugate(...) feeds the Gram contraction only, reference gauge2d.py:1974-1992)
        lz0 = bt.pending()
        if lz0 is None:
            root, pm, conj0 = bt, list(range(n)), False
            A0, B0, Q0 = set(), set(), set()
        else:
            root, pm, conj0 = lz0.src, lz0.perm, lz0.conj
            A0, B0, Q0 = set(lz0.alpha), set(lz0.beta), set(lz0.Q)
        lz = LazyPermute(root, [pm[a] for a in new_order], {pm[a] for a in alpha} ^ A0, {pm[a] for a in beta} ^ B0,
                         {frozenset(pm[a] for a in pr) for pr in Q} ^ Q0, not conj0, mat)
        return _attach_lazy(res, lz)
    res._buf = mat(_src_buf(bt), 1.0)
    if this_fmt == "matrix":
        res = bt_switch_format(res)
    return res


def power_bt(bt, p, rcond=1e-10):
    """element-wise power in matrix format with the |x| > rcond mask: reference power_ds
    (__init__.py:6059-6069) and the *fixed* power_block (:6071-6082, see oracle/ref_harness.py)."""
    this_fmt = bt.fmt
    m = bt if bt.fmt == "matrix" else bt_switch_format(bt)
    if m is bt:
        m = bt.clone()
    check(lib.gtn_pow_rcond(_ptr(m.buf), m.buf.numel(), dtype_code(m.dtype), float(p), float(rcond), _stream()),
          "gtn_pow_rcond")
    count()
    return m if this_fmt == "matrix" else bt_switch_format(m)


# ------------------------------------------------------------------------------------------------
#  user-level join / split of block tensors (reference join_legs_block / split_legs_block)
# ------------------------------------------------------------------------------------------------
def _joined_sigma_bits(lay):
    """sigma exponent bits of a joined fermionic leg, per total parity: sub-block v carries
    prod_{i<j} (-1)^{v_i v_j} * outer(member sigma) (reference join_index :3357-3366)."""
    bits = {0: np.zeros(lay.even_total, dtype=np.uint32), 1: np.zeros(lay.total - lay.even_total, dtype=np.uint32)}
    for v in lay.pats:
        par = sum(v) % 2
        b = np.zeros(1, dtype=np.uint32)
        for pi, d in zip(v, lay.shape[v]):
            b = (b[:, None] ^ sigma_bits(pi, d)[None, :]).ravel()
        perm = sum(v[i] * v[j] for i in range(len(v)) for j in range(i + 1, len(v))) & 1
        o = lay.offset[v] - lay.sector(par)[0]
        bits[par][o: o + lay.size[v]] = b ^ np.uint32(perm)
    return bits[0], bits[1]


def _check_block_groups(groups, stats, final_stat, fn):
    if "*" in tuple(final_stat):
        _err("Error[%s]: Hybrid joining is not allowed for block format." % fn)
    if len(final_stat) != len(groups):
        _err("Error[%s]: Inconsistent number of final statistics and groupings!" % fn)
    for ax, fs in zip(groups, final_stat):
        st = [stats[a] for a in ax]
        f = [s in FERMI for s in st]
        if any(f) and not all(f):
            _err("Error[%s]: Hybrid joining is not allowed for block format." % fn)
        if (fs in FERMI) != all(f):
            _err("Error[%s]: The final statistics are not consistent with the groups." % fn)


def join_block_bt(bt, groups, final_stat):
    """reference join_legs_block (__init__.py:3376-3621): ONE sign+permute launch.  Every parity block goes, as a
    (prod of members) sub-block per group, into block (total parities) at the offset of its sub-block in the
    reference's enumeration order, times (-1)^p for every +1 member of a group whose final statistic is -1
    (:3574-3580); the data carries no sigma factors, the format is kept.
    Returns (joined BT, {joined axis: (sigma bits even, sigma bits odd)})."""
    _check_block_groups(groups, bt.stats, final_stat, "join_legs_block")
    lays = [group_layout([bt.leg(a) for a in ax], "ref") for ax in groups]
    fg = [bt.stats[ax[0]] in FERMI for ax in groups]
    ne = [lay.even_total if f else lay.total for lay, f in zip(lays, fg)]
    no = [lay.total - lay.even_total if f else 0 for lay, f in zip(lays, fg)]
    out = BT(tuple(final_stat), ne, no, bt.dtype, bt.fmt)
    live = bt.live()

    def vpat(pis, ax):
        return tuple(pis[a] for a in ax if a in pis)

    def opat_of(pat):
        pis = dict(zip(bt.faxes, pat))
        return tuple(sum(vpat(pis, ax)) % 2 for ax, f in zip(groups, fg) if f)
    out.alloc(sorted({opat_of(p) for p in live}), zero=True)
    alpha = {a for k, ax in enumerate(groups) if final_stat[k] == -1 for a in ax if bt.stats[a] == 1}

    def build():
        def place(pat, bshape):
            pis = dict(zip(bt.faxes, pat))
            op = opat_of(pat)
            oblk = _row_strides(out.block_shape(op))
            base, ostr = out.off[op], [0] * bt.ndim
            for k, (ax, lay) in enumerate(zip(groups, lays)):
                v = vpat(pis, ax)
                base += (lay.offset[v] - lay.sector(sum(v) % 2)[0]) * oblk[k]
                for j, a in enumerate(ax):
                    ostr[a] = lay.strides(v)[j] * oblk[k]
            return base, ostr, False
        return PermutePlan(_block_jobs(bt, live, alpha, set(), set(), place))
    _cached(("joinblk", bt.key(), tuple(map(tuple, groups)), tuple(final_stat)), build).run(bt.buf, out.buf)
    sigma = {k: _joined_sigma_bits(lay) for k, (lay, f) in enumerate(zip(lays, fg)) if f}
    return out, sigma


def split_block_bt(bt, sigma, groups, final_stat, e, o):
    """reference split_legs_block (__init__.py:3623-3859): standard format first -- with the joined legs' own sigma
    vectors (:3644) --, every joined block cut into its sub-blocks with the (-1)^p correction undone (:3838-3846),
    one sign+permute launch, then back to the caller's format with the STANDARD sigma of the split legs (:3857)."""
    this_fmt = bt.fmt
    S = bt if bt.fmt == "standard" else bt_switch_format(bt, sigma=sigma)
    _check_block_groups(groups, final_stat, S.stats, "split_legs_block")
    res = BT(tuple(final_stat), e, o, bt.dtype, "standard")
    lays = [group_layout([res.leg(a) for a in ax], "ref") for ax in groups]
    fg = [final_stat[ax[0]] in FERMI for ax in groups]
    for k, (lay, f) in enumerate(zip(lays, fg)):
        want = (lay.even_total, lay.total - lay.even_total) if f else (lay.total, 0)
        if (S.e[k], S.o[k]) != want:
            _err("Error[split_legs_block]: The final shape is not consistent with the joined leg %d." % k)

    def jpat_of(pat):
        pis = dict(zip(res.faxes, pat))
        return tuple(sum(pis[a] for a in ax) % 2 for ax, f in zip(groups, fg) if f)
    live_j = set(S.live())
    pats = [p for p in res.patterns() if jpat_of(p) in live_j and res.block_size(p) > 0]
    res.alloc(pats)

    def build():
        jobs = []
        for p in pats:
            pis = dict(zip(res.faxes, p))
            jp = jpat_of(p)
            iblk = _row_strides(S.block_shape(jp))
            oblk = _row_strides(res.block_shape(p))
            base, const, legs = S.off[jp], 0, [None] * res.ndim
            for k, (ax, lay) in enumerate(zip(groups, lays)):
                v = tuple(pis[a] for a in ax if a in pis)
                base += (lay.offset[v] - lay.sector(sum(v) % 2)[0]) * iblk[k]
                st = lay.strides(v)
                for j, a in enumerate(ax):
                    legs[a] = lin_leg(lay.shape[v][j], st[j] * iblk[k], oblk[a])
                    if S.stats[k] == -1 and final_stat[a] == 1:
                        const ^= pis[a]
            jobs.append(build_job(legs, const=const, in_base=base, out_base=res.off[p]))
        return PermutePlan(jobs)
    key = ("splitblk", S.key(), tuple(map(tuple, groups)), tuple(final_stat), tuple(res.e), tuple(res.o))
    _cached(key, build).run(S.buf, res.buf)
    return res if this_fmt == "standard" else bt_switch_format(res)
