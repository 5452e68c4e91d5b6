"""One coarse-graining chain sharded over the GPUs of a box (one process per GPU, torch.distributed; NCCL over
NVLink on the B200 box, gloo for the CPU tests of the host logic).  SURVEY.md section 8(e); BASELINE.json north_star:
"the large-chi contraction is sharded by output tile across the 8 GPUs of one box, with NCCL used only to all-gather
the isometries".

What lives where.  The site tensor T[i,j,k,l] stays SHARDED along its first leg for the whole run: rank r holds
T[i in R_r, :, :, :] as an ordinary parity-blocked tensor whose first leg is shorter (per parity sector, R_r is a
contiguous range of the even and of the odd block indices).  Everything a TRG step does with it keeps that split:

  * T1 = T[jkli], T2 = T[klij] (sign+permute, local).  In both matricisations (jk|li) and (kl|ij) leg i is a COLUMN
    leg, so rank r holds a column subset W[:, C_r] of every parity-sector matrix.
  * truncated sector SVD on column-sharded matrices (ShardedTruncPlan): the subspace iteration of _engine._TruncPlan
    with the sums over columns completed by all-reduces of small panels --
        Yh = G W^H   (l x p)   = sum_r G[:, C_r] W[:, C_r]^H        all-reduce, then orthonormalised on every rank
        Zh = Qh W    (l x q_r)   local columns;  its l x l Gram matrix = sum_r Zh_r Zh_r^H      all-reduce
        B  = Qh W    (l x q_r)   local;  the l x q rows are all-gathered, the one-sided Jacobi SVD of problem b runs
                                 on rank b % W (one owner per problem: W problems in flight instead of W replicas of each) and
                                 Ub, s, Vh are broadcast; every rank keeps its columns of Vh
        certificate rows  Vk W^H - S Uh  (l x p): partial sums over columns, all-reduce
    U (p x l) comes out replicated, V (l x q_r) sharded like the input.  Per iteration 2 x l x (p + l) numbers cross
    NVLink per sector (16 MiB at chi = 128) against 8 p q_r l flops of local GEMM work.
  * the isometries: V1, V2 are all-gathered along the sharded leg (the north-star's all-gather; l x D^2 numbers each),
    VV = V1 . V2 is built on every rank (O(chi^2 D^3), 1/chi of the step's flops).
  * the big contraction T'[i,j,k,l] = sum UU[j,z,x,i] VV[l,x,z,k] is sharded by OUTPUT rows: rank r slices the NEW
    bond index i of U1 (a replicated isometry) to its range, builds UU[.., i in R_r] and contracts it with VV --
    an ordinary local einsum whose result IS rank r's shard of T'.  No collective touches T or T'.
  * Tnorm: local sum of squares, all-reduce of one double.

Not sharded (replicated on every rank): the l x l whitening (one CTA per matrix), the small einsums with the singular
values, VV.  The reference has no distributed mode at all.
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi, _engine as E
from ._cabi import check, count, lib
from ._engine import BT, FERMI, PermutePlan, _cached, _ptr, _row_strides, _stream, build_job, dtype_code, lin_leg

# fused GEMM + all-reduce over peer memory (ShardedTruncPlan._gemm_allreduce): measured equal to GEMM + NCCL all-reduce at
# 2 GPUs (195.2 vs 194.9 ms per chi = 128 step) and slower at 8 (71.5 vs 64.7 ms: an all-reduce by peer stores of the
# partial panels moves W - 1 copies per rank, 2.7 GB per step at W = 8, where NCCL reduces inside the NVSwitch) -> off
FUSED_ALLREDUCE = bool(int(__import__("os").environ.get("GTN_FUSED_ALLREDUCE", "0")))
BIG_PREROTATE = bool(int(__import__("os").environ.get("GTN_BIG_PREROTATE", "1")))
BIG_PREROTATE_MIN_L = 80    # (gtn_gram_rotate serves l <= 80 on the single-GPU path)
GATHER_MAX_DIM = 16384     # largest sector (smaller side) the owner-gathered fallback decomposes on one GPU
STATS = {"allreduce_bytes": 0, "allgather_bytes": 0, "broadcast_bytes": 0, "collectives": 0}


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _real(t):
    return torch.view_as_real(t) if t.is_complex() else t


def _nbytes(t):
    return t.numel() * t.element_size()


# ---- collectives: the library's own NCCL wrappers (include/gtn_b200.h gtn_comm_*, csrc/gtn_comm.cu) on the caller's
#      stream when the process group runs on NCCL; torch.distributed otherwise (gloo: the CPU tests of the host logic)
NATIVE_COMM = bool(int(__import__("os").environ.get("GTN_NATIVE_COMM", "1")))
_native = {"comm": None, "tried": False}
_peer_flags = {}           # symmetric buffer size -> [peer pointers of the barrier flags, epoch]


def native_comm():
    """the C-ABI communicator of this process (created on first use: rank 0's unique id travels through the existing
    torch.distributed group, the host's own means), or None when the group is not on NCCL / GPUs"""
    if _native["tried"]:
        return _native["comm"]
    _native["tried"] = True
    if not (NATIVE_COMM and world() > 1 and dist.get_backend() == "nccl" and lib.gtn_comm_available()):
        return None
    import ctypes as C
    dev = torch.device("cuda", torch.cuda.current_device())
    idbuf = (C.c_char * 128)()
    if rank() == 0:
        check(lib.gtn_comm_unique_id(idbuf), "gtn_comm_unique_id")
    t = torch.frombuffer(bytearray(bytes(idbuf)), dtype=torch.uint8).to(dev)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().numpy().tobytes())
    comm = C.c_void_p()
    check(lib.gtn_comm_init(raw, rank(), world(), C.byref(comm)), "gtn_comm_init")
    _native["comm"] = comm
    return comm


def all_reduce_(t, op=0):
    """in-place sum (op 0) / max (1) / min (2) over the ranks"""
    if world() > 1:
        comm = native_comm()
        with E.prof_region("nccl_allreduce", 0, _nbytes(t)):
            if comm is not None and t.dtype in (torch.float64, torch.complex128) and t.is_contiguous():
                check(lib.gtn_allreduce(comm, _ptr(t), t.numel(), dtype_code(t.dtype), op, _stream()), "gtn_allreduce")
            else:
                dist.all_reduce(_real(t), op=(dist.ReduceOp.SUM, dist.ReduceOp.MAX, dist.ReduceOp.MIN)[op])
        STATS["allreduce_bytes"] += _nbytes(t)
        STATS["collectives"] += 1
    return t


def _all_gather(out, mine):
    comm = native_comm()
    with E.prof_region("nccl_allgather", 0, _nbytes(out)):
        if comm is not None and mine.dtype in (torch.float64, torch.complex128) and mine.is_contiguous() \
                and out.is_contiguous():
            check(lib.gtn_allgather(comm, _ptr(mine), _ptr(out), mine.numel(), dtype_code(mine.dtype), _stream()),
                  "gtn_allgather")
        else:
            dist.all_gather_into_tensor(_real(out), _real(mine))
    STATS["allgather_bytes"] += _nbytes(out)
    STATS["collectives"] += 1


def _broadcast(t, src):
    comm = native_comm()
    with E.prof_region("nccl_broadcast", 0, _nbytes(t)):
        if comm is not None and t.dtype in (torch.float64, torch.complex128) and t.is_contiguous():
            check(lib.gtn_broadcast(comm, _ptr(t), t.numel(), dtype_code(t.dtype), src, _stream()), "gtn_broadcast")
        else:
            dist.broadcast(_real(t), src=src)
    STATS["broadcast_bytes"] += _nbytes(t)
    STATS["collectives"] += 1


def leg_range(n, r, w):
    """[lo, hi) of a sector of extent n owned by rank r of w: contiguous, sizes differ by at most one"""
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


# ------------------------------------------------------------------------------------------------
#  slicing / gathering a parity-blocked tensor along one leg (one launch of the permute kernel each)
# ------------------------------------------------------------------------------------------------
def slice_leg(bt, leg, r=None, w=None):
    """rank r's part of `bt` along `leg`: the even and the odd sector of the leg are cut into w contiguous ranges."""
    r = rank() if r is None else r
    w = world() if w is None else w
    fer = bt.stats[leg] in FERMI
    lo_e, hi_e = leg_range(bt.e[leg], r, w)
    lo_o, hi_o = leg_range(bt.o[leg], r, w) if fer else (0, 0)
    e, o = list(bt.e), list(bt.o)
    e[leg], o[leg] = hi_e - lo_e, hi_o - lo_o
    out = BT(bt.stats, e, o, bt.dtype, bt.fmt)
    pats = [p for p in bt.off if p not in bt.zero]
    out.alloc([p for p in pats if out.block_size(p) > 0])
    out.zero = set()

    def build():
        jobs = []
        for p in out.off:
            pis = dict(zip(bt.faxes, p))
            lo = lo_o if (fer and pis[leg] == 1) else lo_e
            ishape, oshape = bt.block_shape(p), out.block_shape(p)
            istr, ostr = _row_strides(ishape), _row_strides(oshape)
            legs = [lin_leg(oshape[a], istr[a], ostr[a]) for a in range(bt.ndim)]
            order = list(range(bt.ndim))
            jobs.append(build_job(legs, in_base=bt.off[p] + lo * istr[leg], out_base=out.off[p], in_order=order,
                                  out_order=order))
        return PermutePlan(jobs)
    _cached(("slice_leg", bt.key(), leg, r, w), build).run(bt.buf, out.buf)
    return out


def gather_leg(bt, leg, full_e, full_o):
    """inverse of slice_leg on every rank: all-gather of the local buffers (ONE collective), then one launch that places
    every rank's blocks at its range of the full leg.  full_e / full_o: extents of the leg in the full tensor."""
    w, r = world(), rank()
    fer = bt.stats[leg] in FERMI
    e, o = list(bt.e), list(bt.o)
    e[leg], o[leg] = full_e, (full_o if fer else 0)
    out = BT(bt.stats, e, o, bt.dtype, bt.fmt)
    pats = [p for p in bt.off if p not in bt.zero]            # same pattern set on every rank (layouts are data independent)
    out.alloc([p for p in pats if out.block_size(p) > 0])
    # local layouts of all ranks (extents differ by at most one between ranks)
    lays = []
    for q in range(w):
        le, lo_ = leg_range(full_e, q, w), (leg_range(full_o, q, w) if fer else (0, 0))
        eq, oq = list(bt.e), list(bt.o)
        eq[leg], oq[leg] = le[1] - le[0], lo_[1] - lo_[0]
        lq = BT(bt.stats, eq, oq, bt.dtype, bt.fmt)
        acc = 0
        for p in pats:
            if lq.block_size(p) > 0:
                lq.off[p] = acc
                acc += lq.block_size(p)
        lays.append((lq, le[0], lo_[0], acc))
    pad = max(x[3] for x in lays)
    mine_lay = lays[r][0]
    if all(bt.off[p] == o_ for p, o_ in mine_lay.off.items()):
        mine = bt.buf[:lays[r][3]]                              # the live blocks already sit back to back
    else:
        mine = torch.cat([bt.buf[bt.off[p]: bt.off[p] + mine_lay.block_size(p)] for p in mine_lay.off])
    if w == 1:
        gathered = mine
    else:
        if lays[r][3] < pad:
            mine = torch.cat([mine, torch.zeros(pad - lays[r][3], dtype=bt.dtype, device=bt.buf.device)])
        gathered = torch.empty(w * pad, dtype=bt.dtype, device=bt.buf.device)
        _all_gather(gathered, mine.contiguous())

    def build():
        jobs = []
        for q, (lq, lo_e, lo_o, _) in enumerate(lays):
            for p in lq.off:
                pis = dict(zip(bt.faxes, p))
                lo = lo_o if (fer and pis[leg] == 1) else lo_e
                ishape, oshape = lq.block_shape(p), out.block_shape(p)
                istr, ostr = _row_strides(ishape), _row_strides(oshape)
                legs = [lin_leg(ishape[a], istr[a], ostr[a]) for a in range(bt.ndim)]
                order = list(range(bt.ndim))
                jobs.append(build_job(legs, in_base=q * pad + lq.off[p], out_base=out.off[p] + lo * ostr[leg],
                                      in_order=order, out_order=order))
        return PermutePlan(jobs)
    _cached(("gather_leg", bt.key(), leg, full_e, full_o, w), build).run(gathered, out.buf)
    return out


# ------------------------------------------------------------------------------------------------
#  truncated SVD of column-sharded sector matrices
# ------------------------------------------------------------------------------------------------
class ShardedTruncPlan(E._TruncPlan):
    """_engine._TruncPlan on the local column blocks W[:, C_r] (Q_ = local column counts), with the sums over the
    column index completed across ranks.  Every replicated quantity (Yh, Qh, Gram matrices, Ub, s, certificate) is
    bit-identical on all ranks: it is either the result of an all-reduce / broadcast or computed from such results
    by deterministic kernels."""

    def __init__(self, P_, Q_, ks, L_, dt, dev):
        super().__init__(P_, Q_, ks, L_, dt, dev)
        self.prerotate = False                       # (the Gram pre-rotation targets chi <= 32; not used when sharded)
        w = self.w = world()
        # the sketch matrix G (l x q): every rank holds ITS columns, drawn from its own stream (identical blocks on
        # all ranks would project onto the sum of the column blocks only)
        gen = torch.Generator(device="cpu")
        for b in range(self.nb):
            n = L_[b] * Q_[b]
            gen.manual_seed(20240607 + 7919 * rank() + 104729 * b + n)
            if dt == torch.complex128:
                g_ = torch.view_as_complex(torch.randn(n, 2, generator=gen, dtype=torch.float64))
            else:
                g_ = torch.randn(n, generator=gen, dtype=torch.float64)
            self.ws.view(self.hG[b]).view(-1).copy_(g_.to(dev))
        self.Qfull = [q * w for q in Q_]             # equal column counts on every rank (checked by the caller)
        self.gbuf = [torch.empty(w * l * q, dtype=dt, device=dev) for l, q in zip(L_, Q_)]
        # owner-side Jacobi workspace: full rows of B and of Vh, Z, Ub per problem
        jws = self.jws = E._WS(dt, dev)
        self.jB = [jws.add(l, q) for l, q in zip(L_, self.Qfull)]
        self.jV = [jws.add(l, q) for l, q in zip(L_, self.Qfull)]
        self.jZ = [jws.add(l, l) for l in L_]
        self.jU = [jws.add(l, l) for l in L_]
        jws.alloc()
        self.jvh_delta = jws.off(self.jV[0]) - jws.off(self.jB[0])
        self.mine = [b for b in range(self.nb) if b % w == rank()]
        nm = len(self.mine)
        if nm:
            parr, oarr = (_cabi.SvdProblem * nm)(), (_cabi.SvdOut * nm)()
            soff = 0
            self.jsoff = []
            for k, b in enumerate(self.mine):
                parr[k].w_off, parr[k].z_off, parr[k].p, parr[k].q = jws.off(self.jB[b]), jws.off(self.jZ[b]), L_[b], self.Qfull[b]
                oarr[k].s_off, oarr[k].u_off = soff, jws.off(self.jU[b])
                self.jsoff.append(soff)
                soff += L_[b]
            assert all(jws.off(self.jV[b]) - jws.off(self.jB[b]) == self.jvh_delta for b in self.mine)
            self.jpdev, self.jodev = E._to_dev_bytes(bytes(parr)), E._to_dev_bytes(bytes(oarr))
            i64 = lambda v: torch.tensor(list(v), dtype=torch.int64).to(dev)
            self.jrn_off = i64(self.jsoff)
            self.jrn2 = torch.empty(soff, dtype=torch.float64, device=dev)
            self.jfro2 = torch.empty(2 * nm, dtype=torch.float64, device=dev)
            self.joffd = torch.zeros(2 * nm, dtype=torch.float64, device=dev)
            self.jsw = torch.zeros(4, dtype=torch.int32, device=dev)
            self.js = torch.empty(soff, dtype=torch.float64, device=dev)
            self.jorder = torch.empty(soff, dtype=torch.int32, device=dev)
            self.jscratch = torch.empty(soff, dtype=torch.float64, device=dev)
            self.jmaxL = max(L_[b] for b in self.mine)
            self.jmaxQ = max(self.Qfull[b] for b in self.mine)
        self.jacobi_ok = torch.ones(1, dtype=torch.float64, device=dev)
        self._fused_ok = None                        # fused GEMM + all-reduce over peer memory: None = not tried yet
        self.graphable = False                       # collectives between the launches: eager schedule
        torch.cuda.current_stream().synchronize() if dev.type == "cuda" else None

    _allreduce = staticmethod(lambda t: all_reduce_(t))      # rank certificate: sums over the column blocks

    # ---- launch sequences --------------------------------------------------------------------------
    def _reduce_handles(self, handles):
        """all-reduce the (consecutive) workspace matrices `handles` in one collective"""
        ws = self.ws
        a0 = ws.off(handles[0])
        a1 = ws.off(handles[-1]) + ws.items[handles[-1]][1] * ws.items[handles[-1]][2]
        all_reduce_(ws.buf[a0:a1])

    def _gemm_allreduce(self, ha, hb, hc):
        """hc_b = sum over ranks of ha_b . hb_b (the l x p panels whose contracted index is the sharded column index).
        Fused path (NCCL group, <= 8 ranks, peer access): ONE grouped GEMM launch whose epilogue stores this rank's
        partial panels into slot `rank` of EVERY rank's staging buffer over NVLink (gtn_grouped_gemm_bcast on torch
        symmetric memory: the stores overlap the DMMA work of the other resident CTAs), a device-side barrier, and
        gtn_sum_slices adding the W slots in rank order -- deterministic and bit-identical on all ranks, which an
        atomics-based reduction would not be.  Otherwise: local GEMM, then an NCCL all-reduce of the panels."""
        ws, w, r = self.ws, self.w, rank()
        if w > 1 and FUSED_ALLREDUCE and self._fused_ok is not False:
            from . import parallel
            span = ws.off(hc[-1]) + ws.items[hc[-1]][1] * ws.items[hc[-1]][2] - ws.off(hc[0])
            data_bytes = w * span * ws.buf.element_size()
            nbytes = data_bytes + 4096               # + the flag arrays of the device-side barrier
            if self._fused_ok is None:
                err = None
                try:
                    ok_ = dist.get_backend() == "nccl" and w <= parallel.MAX_PEERS
                    if ok_:
                        parallel._symm_output(nbytes, ws.buf.device)
                except Exception as exc:             # no peer access / symmetric memory on this system
                    ok_, err = False, exc
                flag = torch.tensor([1.0 if ok_ else 0.0], dtype=torch.float64, device=ws.buf.device)
                all_reduce_(flag, op=2)              # all ranks take the same path
                self._fused_ok = bool(flag.item() > 0)
                if not self._fused_ok:
                    parallel._symm.pop(nbytes, None)
            if self._fused_ok:
                t, hdl, ptrs = parallel._symm_output(nbytes, ws.buf.device)
                bar = _peer_flags.get(nbytes)
                if bar is None:
                    import ctypes as C
                    t[data_bytes:].zero_()
                    hdl.barrier(channel=0)           # (once per buffer: the flags are zero everywhere)
                    bar = _peer_flags[nbytes] = [(C.c_void_p * w)(*[int(p_) + data_bytes for p_ in hdl.buffer_ptrs]), 0]

                def barrier():
                    bar[1] += 1
                    check(lib.gtn_peer_barrier(bar[0], r, w, bar[1], _stream()), "gtn_peer_barrier")
                    count()
                groups = []
                for a, b, c in zip(ha, hb, hc):
                    _, m, k = ws.items[a]
                    _, _, n = ws.items[b]
                    groups.append(dict(a_off=ws.off(a), b_off=ws.off(b), c_off=r * span + ws.off(c) - ws.off(hc[0]),
                                       lda=k, ldb=n, ldc=n, m=m, n=n, k=k))
                key = ("shard_fused", str(ws.dtype), r, w, tuple(tuple(sorted(g.items())) for g in groups))
                plan = _cached(key, lambda: E.GemmPlan(groups, ws.dtype, config=0))
                barrier()                            # every peer has summed the previous contents of its buffer
                with E.prof_region("grouped_gemm_bcast", 1, plan.bytes, plan.flops):
                    check(lib.gtn_grouped_gemm_bcast(_ptr(ws.buf), _ptr(ws.buf), ptrs, w, dtype_code(ws.dtype),
                                                     _ptr(plan.dev), plan.n, plan.tiles, _stream()),
                          "gtn_grouped_gemm_bcast")
                barrier()                            # every rank's partial panels have landed here
                dst = ws.buf[ws.off(hc[0]): ws.off(hc[0]) + span]
                with E.prof_region("sum_slices", 1, (w + 1) * span * ws.buf.element_size()):
                    check(lib.gtn_sum_slices(_ptr(t), _ptr(dst), span, w, dtype_code(ws.dtype), _stream()),
                          "gtn_sum_slices")
                STATS["fused_allreduce_bytes"] = STATS.get("fused_allreduce_bytes", 0) + (w - 1) * span * ws.buf.element_size()
                STATS["fused_allreduces"] = STATS.get("fused_allreduces", 0) + 1
                return
        E._ws_gemm(ws, list(zip(ha, hb, hc)))
        self._reduce_handles(hc)

    def start(self, passes, robust=False):
        self._gemm_allreduce(self.hG, self.hWh, self.hYh)
        self.orth(self.hYh, self.hQh, "p", passes, robust)

    def iterate(self, last, robust=False):
        ws = self.ws
        E._ws_gemm(ws, list(zip(self.hQh, self.hW, self.hZh)))
        self.orth(self.hZh, self.hPh, "q", 1, robust)
        self._gemm_allreduce(self.hPh, self.hWh, self.hYh)
        self.orth(self.hYh, self.hQh, "p", 2 if last else 1, robust)

    def _gram_done(self, side):
        if side == "q":
            self._reduce_handles(self.hT1)            # Gram matrix of column-sharded rows: sum over ranks

    def check_enqueue(self, allow_host=True):
        ws, nb, dt, dev = self.ws, self.nb, self.dt, self.dev
        code = dtype_code(dt)
        st = _stream()
        w, r = self.w, rank()
        E._ws_gemm(ws, list(zip(self.hQh, self.hW, self.hB)))               # B = Qh W, local columns
        # ---- rows of B to their owners (all-gather; the owner re-packs [rank][row][col] -> [row][rank, col])
        for b in range(nb):
            loc = ws.view(self.hB[b])
            if w > 1:
                _all_gather(self.gbuf[b], loc.reshape(-1))
            else:
                self.gbuf[b].copy_(loc.reshape(-1))
        jws = self.jws
        for b in self.mine:
            l, q = self.L_[b], self.Q_[b]
            jws.view(self.jB[b]).view(l, w, q).copy_(self.gbuf[b].view(w, l, q).permute(1, 0, 2))
        self.jacobi_ok.fill_(1.0)
        rot = {}
        if self.mine and BIG_PREROTATE and min(self.L_[b] for b in self.mine) > BIG_PREROTATE_MIN_L:
            # Pre-rotation for wide subspaces (l > 80: chi >= 128, where gtn_gram_rotate's shared-memory Jacobi does not
            # fit): the Jacobi kernels diagonalise the l x l Gram matrix G = B B^H first (cheap: l^2 numbers per round
            # instead of l q), B' = U_G^H B then has rows orthogonal to the accuracy a Gram matrix allows and the Jacobi
            # SVD of the l x q matrix needs 2 sweeps instead of 8 (12 -> 5 ms per check at chi = 128, the largest
            # unsharded piece of a sharded step); B = U_G B' = (U_G Ub') S Vh: the singular values never go through G.
            gs = []
            for b in self.mine:
                l, Qf = self.L_[b], self.Qfull[b]
                Bm = jws.view(self.jB[b])
                Bh = torch.empty(Qf * l, dtype=dt, device=dev)
                E._ctranspose_t(Bm, l, Qf, Qf, Bh)
                G = torch.empty(l * l, dtype=dt, device=dev)
                E._gemm_t(Bm, Qf, Bh, l, G, l, l, l, Qf)
                gs.append(G.view(l, l))
            for b, (Ug, _, _) in zip(self.mine, E.batched_svd(gs)):
                l, Qf = self.L_[b], self.Qfull[b]
                Ug = Ug.contiguous()
                Th = torch.empty(l * l, dtype=dt, device=dev)
                E._ctranspose_t(Ug, l, l, l, Th)                                # T = U_G^H
                Bm = jws.view(self.jB[b])
                Bp = torch.empty(l * Qf, dtype=dt, device=dev)
                E._gemm_t(Th, l, Bm, Qf, Bp, Qf, l, Qf, l)
                Bm.view(-1).copy_(Bp)
                rot[b] = Ug
        if self.mine:
            nm = len(self.mine)
            Wp = _ptr(jws.buf)
            check(lib.gtn_jacobi_init(Wp, Wp, code, _ptr(self.jpdev), nm, self.jmaxL, _ptr(self.jrn2), _ptr(self.jfro2),
                                      _ptr(self.jrn_off), st), "gtn_jacobi_init")
            count()
            done = False
            P = (self.jmaxL + 1) & ~1
            esz = jws.buf.element_size()
            rb = sum(2 * esz * (self.L_[b] * self.Qfull[b] + self.L_[b] ** 2) for b in self.mine)
            if 2 <= self.jmaxL <= E.PERSISTENT_MAX_ROWS:
                with E.prof_region("jacobi_persistent", 1, 0):
                    rc = lib.gtn_jacobi_persistent(Wp, Wp, code, _ptr(self.jpdev), nm, self.jmaxL, E.JACOBI_TOL,
                                                   _ptr(self.joffd), _ptr(self.jrn2), _ptr(self.jfro2), _ptr(self.jrn_off),
                                                   E.JACOBI_MAX_SWEEPS, _ptr(self.jsw), st, E.JACOBI_EARLY_STOP)
                if rc == 0:
                    done = True
                    self.jacobi_ok.copy_(self.jsw[1:2].to(torch.float64))       # converged flag, read with the certificate
                elif rc != -2:
                    check(rc, "gtn_jacobi_persistent")
            if not done:
                sweeps = 0
                self.joffd.zero_()
                while True:
                    with E.prof_region("jacobi_round", P - 1, rb * (P - 1)):
                        check(lib.gtn_jacobi_sweep(Wp, Wp, code, _ptr(self.jpdev), nm, self.jmaxL, self.jmaxQ, E.JACOBI_TOL,
                                                   _ptr(self.joffd), _ptr(self.jrn2), _ptr(self.jfro2), _ptr(self.jrn_off), st),
                              "gtn_jacobi_sweep")
                    sweeps += 1
                    if float(self.joffd[:nm].max().item()) <= E.JACOBI_TOL ** 2:
                        break
                    if sweeps >= E.JACOBI_MAX_SWEEPS:
                        self.jacobi_ok.fill_(0.0)
                        break
            import ctypes as C
            vh_ptr = C.c_void_p(jws.buf.data_ptr() + self.jvh_delta * esz)
            check(lib.gtn_jacobi_finish(Wp, Wp, Wp, vh_ptr, _ptr(self.js), code, _ptr(self.jpdev), _ptr(self.jodev),
                                        _ptr(self.jorder), _ptr(self.jscratch), nm, self.jmaxL, self.jmaxQ, st),
                  "gtn_jacobi_finish")
            count(2)
            for k, b in enumerate(self.mine):
                self.s_dev[self.soff[b]: self.soff[b] + self.L_[b]].copy_(self.js[self.jsoff[k]: self.jsoff[k] + self.L_[b]])
                if b in rot:                                                    # Ub = U_G Ub'
                    l = self.L_[b]
                    Ub = jws.view(self.jU[b])
                    tmp = torch.empty(l * l, dtype=dt, device=dev)
                    E._gemm_t(rot[b], l, Ub, l, tmp, l, l, l, l)
                    Ub.view(-1).copy_(tmp)
        # ---- Ub, s, Vh from the owners
        if w > 1:
            all_reduce_(self.jacobi_ok, op=2)
            for b in range(nb):
                src = b % w
                for t in (jws.view(self.jU[b]), jws.view(self.jV[b]), self.s_dev[self.soff[b]: self.soff[b] + self.L_[b]]):
                    _broadcast(t, src)
        for b in range(nb):
            l, q = self.L_[b], self.Q_[b]
            ws.view(self.hUb[b]).copy_(jws.view(self.jU[b]))
            ws.view(self.hVk[b]).copy_(jws.view(self.jV[b]).view(l, w, q)[:, r, :])
        self.host_sweeps = 0
        E._ws_ctranspose(ws, list(zip(self.hUb, self.hUbH)))
        E._ws_gemm(ws, list(zip(self.hUbH, self.hQh, self.hUh)))            # Uh = Ub^H Qh  (l x p), replicated
        # certificate rows  Eh_i = v_i^H W^H - s_i u_i^H  (l x p): partial sums over the local columns, then all-reduce
        self._gemm_allreduce(self.hVk, self.hWh, self.hXh)
        for b in range(nb):
            ws.view(self.hD[b]).diagonal().copy_(self.s_dev[self.soff[b]: self.soff[b] + self.L_[b]])
        E._ws_gemm(ws, list(zip(self.hD, self.hUh, self.hXh)), alpha=-1.0, beta=1.0)
        for b in range(nb):
            x = ws.view(self.hXh[b])
            with E.prof_region("row_sumsq", 1, x.numel() * x.element_size()):
                check(lib.gtn_row_sumsq(_ptr(x), _ptr(self.res2[self.soff[b]:]), self.L_[b], self.P_[b], code, st),
                      "gtn_row_sumsq")
        sL = self.sumL
        self.out_dev[:sL].copy_(self.s_dev)
        self.out_dev[sL: 2 * sL].copy_(self.res2)
        self.out_dev[2 * sL: 2 * sL + nb].copy_(self.kept[0])
        self.out_dev[2 * sL + nb].fill_(0.0)
        self.out_dev[2 * sL + nb + 1: 2 * sL + nb + 2].copy_(self.jacobi_ok)
        if w > 1:
            _broadcast(self.out_dev, 0)                 # one decision for all ranks, whatever the last bits say
        self.out_host.copy_(self.out_dev, non_blocking=True)

    def read(self):
        if self.dev.type == "cuda":
            torch.cuda.current_stream().synchronize()
        o = self.out_host.numpy()
        sL, nb = self.sumL, self.nb
        svals = [o[a: a + l].copy() for a, l in zip(self.soff, self.L_)]
        res = np.sqrt(np.maximum(o[sL: 2 * sL], 0.0))
        kept = o[2 * sL: 2 * sL + nb].astype(np.int64)
        if not int(o[2 * sL + nb + 1]):
            raise _cabi.GtnError("Jacobi SVD of a projected matrix did not converge on its owner rank")
        return svals, res, kept


_plans = {}


def truncated_svd_sharded(mats, ks, site=None, robust=False):
    """Top-k_b triplets of column-sharded matrices: mats[b] = W_b[:, C_r] on rank r (same shapes on every rank).
    Returns [(U (p x l, replicated), s (host), Vh (l x q_r, local columns))] or None when the certificate cannot be
    met (flat spectrum at the cut) -- the caller then leaves the sharded path for that decomposition.
    robust=True: panels orthonormalised by the Jacobi kernels instead of the Gram whitening (spectra that span more
    than 3e-7 inside the cut); tried automatically when the plain run is rejected, and remembered per site."""
    dev, dt = mats[0].device, mats[0].dtype
    w = world()
    P_ = [m.shape[0] for m in mats]
    Q_ = [m.shape[1] for m in mats]
    L_ = [min(p, q * w, E.subspace_rows(k), E.TRUNC_LMAX) for p, q, k in zip(P_, Q_, ks)]
    pkey = ("sharded", tuple(P_), tuple(Q_), tuple(ks), str(dt), str(dev), w)
    key = (pkey, site)
    if not robust:
        out = None if E._trunc_robust.get(key) else _truncated_svd_sharded(mats, ks, key, pkey, P_, Q_, L_, False)
        if out is None:
            out = _truncated_svd_sharded(mats, ks, key + ("robust",), pkey, P_, Q_, L_, True)
            if out is not None:
                E._trunc_robust[key] = True
                STATS["robust_svds"] = STATS.get("robust_svds", 0) + 1
        return out
    return _truncated_svd_sharded(mats, ks, key + ("robust",), pkey, P_, Q_, L_, True)


def _truncated_svd_sharded(mats, ks, key, pkey, P_, Q_, L_, robust):
    dev, dt = mats[0].device, mats[0].dtype
    plan = _plans.get(pkey)
    if plan is None:
        if len(_plans) >= 4:
            _plans.pop(next(iter(_plans)))
        plan = _plans[pkey] = ShardedTruncPlan(P_, Q_, ks, L_, dt, dev)
    plan.load(mats)
    hint = E._trunc_iters_hint.get(key)
    start_it = hint or 0
    prev_worst, prev_it, next_check = None, None, 0
    for it in range(E.TRUNC_MAX_ITERS + 1):
        if it == 0:
            plan.start(2 if start_it == 0 else 1, robust)
        else:
            plan.iterate(it >= start_it and it >= next_check, robust)
        if it < start_it or it < next_check:
            continue
        plan.check_enqueue()
        try:
            svals, res, kept_host = plan.read()
        except _cabi.GtnError:
            return None
        if robust:
            kept_host = None
        ok, worst, reject = E._trunc_certificate(svals, res, kept_host, ks, L_, rank_check=plan.rank_certificate)
        E.truncated_svd_batch.last_iters = it
        if E.DEBUG_TRUNC and rank() == 0:
            print("[trunc sharded] it", it, "worst %.2e" % worst, "ok", ok, "reject", reject, "robust", robust, flush=True)
        if reject:
            return None
        if prev_worst is not None and prev_worst > 0 and worst > 0 and it > prev_it:
            E._trunc_rate[key] = min(max((worst / prev_worst) ** (1.0 / (it - prev_it)), 1e-3), 0.9)
        if not ok and prev_worst is not None and it >= 2:
            rate = (worst / prev_worst) ** (1.0 / max(it - prev_it, 1)) if prev_worst > 0 else 1.0
            if rate > 0.6:
                return None
            need = math.log(max(E.TRUNC_TOL * 0.3, 1e-300) / max(worst, 1e-300)) / math.log(max(rate, 1e-3))
            next_check = it + max(1, min(int(math.ceil(need)), 6))
            if next_check > E.TRUNC_MAX_ITERS:
                return None
        if not ok and prev_worst is None and key in E._trunc_rate and worst > 0:
            need = math.log(max(E.TRUNC_TOL * 0.3, 1e-300) / worst) / math.log(E._trunc_rate[key])
            next_check = min(it + max(1, min(int(math.ceil(need)), 6)), E.TRUNC_MAX_ITERS)
        prev_worst, prev_it = worst, it
        if ok:
            E._trunc_accept(key, it, worst, spare=1, clean=(it == start_it))
            out = plan.finalize()
            return [(u, svals[b], v) for b, (u, _, v) in enumerate(out)]
    return None


# ------------------------------------------------------------------------------------------------
#  sharded TRG step
# ------------------------------------------------------------------------------------------------
def shard(T, leg=0):
    """local part of a tensor that is IDENTICAL on every rank (e.g. broadcast from rank 0)"""
    import grassmanntn_b200 as gtn
    bt = T._bt if isinstance(T, gtn.block) else T._get_bt()
    out = gtn.block._from_bt(slice_leg(bt, leg))
    out._shard_full = (bt.e[leg], bt.o[leg])
    return out


def unshard(Tl, full_e=None, full_o=None, leg=0):
    """the full tensor on every rank (all-gather along the sharded leg `leg`)"""
    import grassmanntn_b200 as gtn
    if full_e is None:
        full_e, full_o = Tl._shard_full
    return gtn.block._from_bt(gather_leg(Tl._bt, leg, full_e, full_o))


def broadcast_tensor(T, src=0):
    """make rank `src`'s tensor the tensor of every rank (same layout assumed): one broadcast of the block buffer"""
    if world() > 1:
        _broadcast(T._bt.buf, src)
    return T


def svd_many_sharded(objs, string, cutoff, site=None, gathered_fallback=True):
    """gtn.svd_many for tensors sharded along a COLUMN leg of the partition: U replicated, V sharded like the input.
    When the column-sharded truncated path does not apply or fails its certificate, the sector matrices are gathered on
    their owner ranks and decomposed there (_svd_gathered); with gathered_fallback=False None is returned instead."""
    import grassmanntn_b200 as gtn
    from . import _ops, _planner
    left, right = _planner.split_partition(string, "svd")
    nl = len(left)
    ctxs = [_ops._decompose_prepare(o._bt, nl, "svd") for o in objs]
    mats = [m for c in ctxs for m in c["mats"]]
    ks = _ops._sector_cuts(ctxs, cutoff, "block")
    w = world()
    full_min = [min(m.shape[0], m.shape[1] * w) for m in mats]
    usv = None
    if all(k >= 1 and 3 * k // 2 + 8 <= E.TRUNC_LMAX and 4 * min(E.subspace_rows(k), E.TRUNC_LMAX) <= fm
           for k, fm in zip(ks, full_min)):
        usv = truncated_svd_sharded(mats, ks, site=site)
        if usv is not None:
            _ops.SVD_PATH_STATS["truncated"] += 1
    if usv is None:
        if not gathered_fallback:
            return None
        STATS["gathered_svds"] = STATS.get("gathered_svds", 0) + 1
        usv = _svd_gathered(mats, ks, cutoff, site=site)
    outs, k = [], 0
    for c, o in zip(ctxs, objs):
        n = len(c["mats"])
        U, S, V, _ = _ops._decompose_finish(c, usv[k:k + n], cutoff, "svd", "block")
        outs.append(tuple(gtn.block._from_bt(x) for x in (U, S, V)))
        k += n
    return outs


def _svd_gathered(mats, ks, cutoff, site=None):
    """Fallback for sector matrices whose column-sharded subspace iteration cannot be certified (flat spectrum at the
    cut, numerically rank-deficient sectors, matrices too small for the truncated path): the column blocks of problem
    b are gathered on rank b % W, which runs the single-GPU decomposition (_ops._svd_core: truncated SVD with
    certificate, else the full Jacobi SVD), and the first k triplets are broadcast.  Different problems are solved
    concurrently by different owners.  Returns [(U (p x k), s (host), Vh (k x q_r, local columns))]."""
    from . import _ops
    w, r = world(), rank()
    dev, dt = mats[0].device, mats[0].dtype
    if max(min(m.shape[0], m.shape[1] * w) for m in mats) > GATHER_MAX_DIM:
        import grassmanntn_b200 as gtn
        gtn.error("Error[sharded]: the truncated decomposition of a %d x %d sector did not meet its certificate and the "
                  "sector is too large for the single-GPU fallback (full SVD)."
                  % (mats[0].shape[0], mats[0].shape[1] * w))
    full = {}
    for b, m in enumerate(mats):                                  # column blocks to the owners
        p, q = m.shape
        owner = b % w
        mc = m.contiguous()
        if w == 1:
            full[b] = mc
            continue
        parts = [torch.empty(p, q, dtype=dt, device=dev) for _ in range(w)] if r == owner else None
        with E.prof_region("nccl_gather", 0, _nbytes(mc) * w):
            dist.gather(_real(mc), [_real(t) for t in parts] if parts is not None else None, dst=owner)
        STATS["allgather_bytes"] += _nbytes(mc) * w
        STATS["collectives"] += 1
        if r == owner:
            full[b] = torch.cat(parts, dim=1)                     # columns ordered [rank][local column]
            del parts
    mine = sorted(full)
    res = {}
    if mine:
        E.SVD_SITE[0] = (site, "gathered")
        usv, _ = _ops._svd_core([full[b] for b in mine], [ks[b] for b in mine], cutoff, "svd")
        for j, b in enumerate(mine):
            U, sv, Vh = usv[j]
            kk = min(ks[b], mats[b].shape[0], mats[b].shape[1] * w)
            res[b] = (U[:, :kk].contiguous(), np.asarray(sv[:kk], dtype=np.float64), Vh[:kk, :].contiguous())
    del full
    outs = []
    for b, m in enumerate(mats):                                  # first k triplets from the owners
        p, q = m.shape
        kk = min(ks[b], p, q * w)
        if w == 1:
            outs.append(res[b])
            continue
        owner = b % w
        if r == owner:
            U, sv, Vh = res[b]
            sd = torch.from_numpy(sv).to(dev)
        else:
            U = torch.empty(p, kk, dtype=dt, device=dev)
            Vh = torch.empty(kk, q * w, dtype=dt, device=dev)
            sd = torch.empty(kk, dtype=torch.float64, device=dev)
        for t in (U, Vh, sd):
            _broadcast(t, owner)
        outs.append((U, sd.cpu().numpy(), Vh.view(kk, w, q)[:, r, :].contiguous()))
    return outs


def trg(Tl, dcut, full_e=None, full_o=None):
    """One Levin-Nave TRG step (reference gauge2d_block.py:1649-1755) on a site tensor sharded along its first leg.
    Tl: this rank's shard (gtn.block).  Returns (shard of T', Tnorm).  full_e / full_o: extents of the sharded leg in
    the full tensor (default: local extent x world size)."""
    import grassmanntn_b200 as gtn
    w, r = world(), rank()
    bt = Tl._bt
    if full_e is None:
        full_e, full_o = getattr(Tl, "_shard_full", (bt.e[0] * w, bt.o[0] * w))
    if (bt.e[0], bt.o[0]) != (leg_range(full_e, r, w)[1] - leg_range(full_e, r, w)[0],
                              leg_range(full_o, r, w)[1] - leg_range(full_o, r, w)[0]):
        gtn.error("Error[sharded.trg]: the local extent of the first leg does not match the full extent.")
    if full_e % w or full_o % w:
        gtn.error("Error[sharded.trg]: the even and odd extents of the sharded leg must be multiples of the number "
                  "of ranks (the column blocks of the sector matrices have to be equally wide).")
    T1 = gtn.einsum("ijkl->jkli", Tl)                   # leg i: last column leg
    T2 = gtn.einsum("ijkl->klij", Tl)                   # leg i: first column leg
    res = svd_many_sharded([T1, T2], "ab|cd", dcut, site=("trg", "sharded"))
    if res is None:
        gtn.error("Error[sharded.trg]: the truncated decomposition did not meet its certificate on the sharded "
                  "matrices; gather the tensor (sharded.unshard) and run gauge2d.trg for this step.")
    (U1, S1, V1), (U2, S2, V2) = res
    sq = gtn.sqrt(S1)
    U1 = gtn.einsum("abx,xc->abc", U1, sq)
    V1 = gtn.einsum("ax,xbc->abc", sq, V1)
    sq = gtn.sqrt(S2)
    U2 = gtn.einsum("abx,xc->abc", U2, sq)
    V2 = gtn.einsum("ax,xbc->abc", sq, V2)
    # the isometries' sharded legs (T's leg i: V1[x, l, i], V2[x, i, j]) are all-gathered
    V1 = gtn.block._from_bt(gather_leg(V1._bt, 2, full_e, full_o))
    V2 = gtn.block._from_bt(gather_leg(V2._bt, 1, full_e, full_o))
    VV = gtn.einsum("kwz,lxw->lxzk", V1, V2)
    # output rows: the new leg i is U1's bond index -- this rank takes its range of it
    new_full = (U1._bt.e[2], U1._bt.o[2])
    if new_full[0] % w or new_full[1] % w:
        gtn.error("Error[sharded.trg]: the new bond dimension (%d even + %d odd) does not split evenly over %d ranks."
                  % (new_full + (w,)))
    U1r = gtn.block._from_bt(slice_leg(U1._bt, 2))
    UU = gtn.einsum("yxi,zyj->jzxi", U1r, U2)
    Tn = gtn.einsum("lxzk,jzxi->ijkl", VV, UU)
    acc = Tn._bt.sumsq()
    all_reduce_(acc)
    Tnorm = math.sqrt(float(acc.item()))
    out = Tn * (1.0 / Tnorm)
    out._shard_full = new_full
    return out, Tnorm


# ------------------------------------------------------------------------------------------------
#  sharded ATRG step
# ------------------------------------------------------------------------------------------------
def _swap_xy(T):
    import grassmanntn_b200 as gtn
    return gtn.einsum("jikl->jilk", gtn.einsum("ijkl->jikl", T))


def _need(cond, what):
    if not cond:
        import grassmanntn_b200 as gtn
        gtn.error("Error[sharded.atrg]: %s does not split evenly over %d ranks." % (what, world()))


def _blk(bt):
    import grassmanntn_b200 as gtn
    return gtn.block._from_bt(bt)


def atrg2dy(Tl, dcut, intermediate_dcut=None):
    """One ATRG step along y (reference gauge2d_block.py atrg2dy, gauge2d.py:1761-1869, T1 is T2) on a site tensor
    T[i,j,k,l] sharded along its SECOND leg j.  Returns (T' sharded along its FIRST leg, Tnorm): a y step consumes the
    legs i, k and hands j, l through, the x step that follows (atrg2dx) consumes j, l -- so the result is sharded on
    a new bond, which costs nothing because the isometries it is built from are replicated.

    Where the sharded leg sits in the three decompositions (always a COLUMN leg, so every sector matrix is a column
    block W[:, C_r] for sharded.truncated_svd_sharded):
        T[l,i | j,k]  (j)    ->  V1[a,j,k] sharded on j, U1 replicated
        M[a,i | b,k]         <-  C = S1 V1 is all-gathered (an isometry: chi D^2 numbers) and re-cut along k, so that
                                 M = C . B comes out sharded on k without a reduction over the sharded j
        Q[i,j | k,l]  (l)    <-  Q2 = X . A with A = V1 (sharded on T's j leg = Q's l), X all-gathered; Q1 replicated
    and T' = H . G with G all-gathered and H cut along its new bond.  Collectives: three isometry-sized all-gathers
    plus the panels of the three decompositions."""
    import grassmanntn_b200 as gtn
    E_ = gtn.einsum
    w = world()
    chi_i = dcut if intermediate_dcut is None else intermediate_dcut
    bt = Tl._bt
    full_j = getattr(Tl, "_shard_full", (bt.e[1] * w, bt.o[1] * w))
    full_k = (bt.e[2], bt.o[2])
    _need(full_j[0] % w == 0 and full_j[1] % w == 0, "the sharded leg j")
    _need(full_k[0] % w == 0 and full_k[1] % w == 0, "leg k")

    def svd1(obj, string, cut, tag):
        res = svd_many_sharded([obj], string, cut, site=("atrg", "sharded", tag))
        if res is None:
            gtn.error("Error[sharded.atrg]: decomposition %s did not meet its certificate on the sharded matrices; "
                      "gather the tensor (sharded.unshard) and run gauge2d.atrg2dy for this step." % tag)
        return res[0]
    Tr = E_("ijkl->lijk", Tl)                                   # [l,i,j,k], sharded on index 2
    U1, S1, V1 = svd1(Tr, "li|jk", chi_i, 1)
    del Tr
    A = V1                                                       # [a, j in R_r, k]
    B = E_("lia,ab->lib", U1, S1)                                # replicated
    C = E_("ab,bjk->ajk", S1, V1)                                # sharded on j
    C = _blk(gather_leg(C._bt, 1, *full_j))                      # all-gather (isometry) ...
    C = _blk(slice_leg(C._bt, 2))                                # ... and re-cut along k
    M = E_("ajk,jib->aibk", C, B)                                # sharded on index 3 (k)
    U, S, V = svd1(M, "ai|bk", chi_i, 2)
    del M, C, B
    sq = gtn.sqrt(S)
    Y = E_("abx,xc->abc", U, sq)                                 # replicated
    X = E_("ax,xbc->abc", sq, V)                                 # sharded on index 2 (k)
    X = _blk(gather_leg(X._bt, 2, *full_k))
    # (reference: Q1[i,j,a,b]; written here in the GEMM's natural leg order [i a][b j] -- the same Grassmann tensor, one
    #  full pass and one chi^2 D^2-sized temporary fewer: 32 GiB at D = chi = 256 -- and contracted through its labels)
    Q1 = E_("iax,xbj->iabj", U1, Y)                              # replicated (D = U2 = U1)
    Q2 = E_("kya,ylb->abkl", X, A)                               # sharded on index 3 (T's j leg)
    Q = E_("iabj,abkl->ijkl", Q1, Q2)                            # sharded on index 3
    del Q1, Q2
    U, S, V = svd1(Q, "ij|kl", dcut, 3)
    del Q
    sq = gtn.sqrt(S)
    H = E_("abx,xc->abc", U, sq)                                 # replicated
    G = E_("ax,xbc->abc", sq, V)                                 # sharded on index 2 (T's j leg)
    G = _blk(gather_leg(G._bt, 2, *full_j))
    H = E_("lai->ila", H)
    G = E_("kaj->ajk", G)
    new_full = (H._bt.e[0], H._bt.o[0])
    _need(new_full[0] % w == 0 and new_full[1] % w == 0, "the new bond (%d even + %d odd)" % new_full)
    Hr = _blk(slice_leg(H._bt, 0))
    Tn = E_("ila,ajk->ijkl", Hr, G)                              # sharded on index 0 (new bond)
    acc = Tn._bt.sumsq()
    all_reduce_(acc)
    Tnorm = math.sqrt(float(acc.item()))
    out = Tn * (1.0 / Tnorm)
    out._shard_full = new_full
    return out, Tnorm


def atrg2dx(Tl, dcut, intermediate_dcut=None):
    """One ATRG step along x (reference atrg2dx, gauge2d.py:1871-1889: swap the legs, atrg2dy, swap back) on a site
    tensor sharded along its FIRST leg; returns (T' sharded along its SECOND leg, Tnorm) -- ready for atrg2dy."""
    Ts = _swap_xy(Tl)                                            # the sharded leg moves from index 0 to index 1
    Ts._shard_full = Tl._shard_full
    out, Tnorm = atrg2dy(Ts, dcut, intermediate_dcut)
    res = _swap_xy(out)                                          # sharded on index 1 now
    res._shard_full = out._shard_full
    return res, Tnorm
