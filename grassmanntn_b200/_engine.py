"""Device engine: parity-blocked tensor storage + launch builders for the sm_100a kernels.

Everything numeric happens in libgtn_b200.so (csrc/*.cu) through grassmanntn_b200._cabi; this
module only does host-side bookkeeping: which parity block goes where, with which scalar sign,
and which legs carry a sigma vector.  torch is used for device memory and streams only.

Storage (class BT): one contiguous device buffer holding the parity blocks of a Grassmann
tensor, block (pi_1..pi_n) of the fermionic legs being a row-major array over ALL legs with
extent e_a (pi_a = 0) or o_a (pi_a = 1) on fermionic leg a.  Element (pi, s) of a fermionic leg
is parity-preserving index 2s+pi, canonical index encoder(2s+pi) -- the reference's block
format (reference __init__.py:242-344, :690-729).  A dense canonical tensor with power-of-two
fermionic dims is the special case e = o = d/2, and is converted both ways by ONE sign-free
launch of the permute kernel.
"""
import ctypes as C
import itertools
import math

import numpy as np
import torch

from . import _cabi
from ._cabi import AxisEntry, GemmGroup, PermuteJob, SvdOut, SvdProblem, check, count, lib
from .param import _popcount_vec, canonical_of_block, encoder_array

ENTRY_DT = np.dtype([("in_off", "<i8"), ("out_off", "<i8"), ("P", "<u4"), ("M", "<u4")])
assert ENTRY_DT.itemsize == C.sizeof(AxisEntry)
TILE = 32
FERMI = (1, -1)


def _stream():
    # raw handle of torch's current stream (torch.cuda.current_stream() costs ~15 us per call)
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


class _Profiler:
    """Optional per-kernel-family timing with CUDA events on the launching stream (bench.py)."""

    def __init__(self):
        self.enabled = False
        self.records = []
        self.descs = []
        self.detail = []

    def start(self):
        self.enabled = True
        self.records = []
        self.descs = []

    def stop(self, detail=False):
        self.enabled = False
        torch.cuda.synchronize()
        out = {}
        if detail:
            self.detail = [(name, s.elapsed_time(e), flops, nbytes, desc) for (name, s, e, _, nbytes, flops), desc
                           in zip(self.records, self.descs)]
        self.descs = []
        for name, s, e, launches, nbytes, flops in self.records:
            d = out.setdefault(name, dict(ms=0.0, calls=0, launches=0, bytes=0, flops=0))
            d["ms"] += s.elapsed_time(e)
            d["calls"] += 1
            d["launches"] += launches
            d["bytes"] += nbytes
            d["flops"] += flops
        self.records = []
        return out


PROF = _Profiler()


class prof_region:
    def __init__(self, name, launches=1, nbytes=0, flops=0, desc=None):
        self.a = (name, launches, nbytes, flops)
        self.desc = desc

    def __enter__(self):
        if PROF.enabled:
            self.s = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def set_bytes(self, nbytes):
        self.a = (self.a[0], self.a[1], nbytes, self.a[3])

    def __exit__(self, *exc):
        count(self.a[1])
        if PROF.enabled:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            PROF.records.append((self.a[0], self.s, e, self.a[1], self.a[2], self.a[3]))
            PROF.descs.append(self.desc)
        return False


def dtype_code(dt):
    if dt == torch.complex128:
        return _cabi.GTN_C128
    if dt == torch.float64:
        return _cabi.GTN_F64
    raise TypeError("grassmanntn_b200 computes in float64 / complex128 only, got %s" % dt)


_cuda_ok = None


def require_cuda():
    global _cuda_ok
    if _cuda_ok is None:
        _cuda_ok = torch.cuda.is_available()
    if not _cuda_ok:
        _cuda_ok = None
        raise RuntimeError("grassmanntn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _to_dev_bytes(raw):
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(require_cuda(), non_blocking=True)


# ------------------------------------------------------------------------------------------------
#  generic sign+permute launch builder
# ------------------------------------------------------------------------------------------------
class LegTab:
    """Per-leg tables over the leg's index values: element offsets on both sides and the parity /
    sigma bits (None = all zero)."""
    __slots__ = ("n", "in_off", "out_off", "p", "q")

    def __init__(self, n, in_off, out_off, p=None, q=None):
        self.n = int(n)
        self.in_off = np.asarray(in_off, dtype=np.int64)
        self.out_off = np.asarray(out_off, dtype=np.int64)
        self.p = None if p is None else np.asarray(p, dtype=np.uint32)
        self.q = None if q is None else np.asarray(q, dtype=np.uint32)


def lin_leg(n, in_stride, out_stride, q=None, p=None):
    r = np.arange(n, dtype=np.int64)
    return LegTab(n, r * in_stride, r * out_stride, p, q)


def _fuse(legs, ids, alpha, beta, Q):
    """Fuse legs `ids` (slowest first) into one super-axis table."""
    n = 1
    in_off = np.zeros(1, dtype=np.int64)
    out_off = np.zeros(1, dtype=np.int64)
    P = np.zeros(1, dtype=np.uint32)
    intra = np.zeros(1, dtype=np.uint32)
    for a in ids:
        L = legs[a]
        in_off = (in_off[:, None] + L.in_off[None, :]).ravel()
        out_off = (out_off[:, None] + L.out_off[None, :]).ravel()
        pa = L.p if L.p is not None else np.zeros(L.n, dtype=np.uint32)
        qa = L.q if L.q is not None else np.zeros(L.n, dtype=np.uint32)
        ia = np.zeros(L.n, dtype=np.uint32)
        if alpha[a]:
            ia ^= pa
        if beta[a]:
            ia ^= qa
        # cross terms with the legs already fused: popcount(P & Qrow[a]) * p_a
        cross = (_popcount_vec(P & np.uint32(Q[a])) & 1).astype(np.uint32)
        intra = ((intra[:, None] ^ ia[None, :]) ^ (cross[:, None] & pa[None, :])).ravel()
        P = (P[:, None] | (pa[None, :] << np.uint32(a))).ravel()
        n *= L.n
    M = np.zeros(n, dtype=np.uint32)
    for a in ids:
        if Q[a]:
            M ^= np.where((P >> np.uint32(a)) & 1, np.uint32(Q[a]), np.uint32(0)).astype(np.uint32)
    tab = np.empty(n, dtype=ENTRY_DT)
    tab["in_off"] = in_off
    tab["out_off"] = out_off
    tab["P"] = P | (intra << np.uint32(31))
    tab["M"] = M
    return tab


def _choose_groups(dims, in_order, out_order, target=TILE, cap=1 << 16):
    live_in = [a for a in in_order if dims[a] > 1]
    live_out = [a for a in out_order if dims[a] > 1]
    ones = [a for a in range(len(dims)) if dims[a] == 1]
    if not live_in:
        return [list(range(len(dims)))], False
    la, lb = live_in[-1], live_out[-1]
    if la == lb:
        A, size = [la], dims[la]
        k = len(live_in) - 2
        while k >= 0 and size < target and size * dims[live_in[k]] <= cap:
            A.insert(0, live_in[k])
            size *= dims[live_in[k]]
            k -= 1
        rem = [a for a in live_in if a not in A]
        B, size = [], 1
        while rem and size < 8:
            B.insert(0, rem.pop())
            size *= dims[B[0]]
        transpose = False
    else:
        A, B = [la], [lb]
        sa, sb = dims[la], dims[lb]
        k = len(live_in) - 2
        while k >= 0 and sa < target and live_in[k] not in B and sa * dims[live_in[k]] <= cap:
            A.insert(0, live_in[k])
            sa *= dims[live_in[k]]
            k -= 1
        k = len(live_out) - 2
        while k >= 0 and sb < target and live_out[k] not in A and sb * dims[live_out[k]] <= cap:
            B.insert(0, live_out[k])
            sb *= dims[live_out[k]]
            k -= 1
        rem = [a for a in live_in if a not in A and a not in B]
        transpose = True
    A = ones + A
    groups = [A] + ([B] if B else []) + [[a] for a in rem]
    while len(groups) > _cabi.GTN_MAX_SUPER:
        g = groups.pop()
        groups[-1] = groups[-1] + g
    return groups, transpose


def build_job(legs, alpha=None, beta=None, Q=None, const=0, conj=False, in_base=0, out_base=0,
              in_order=None, out_order=None):
    """Returns (job_fields, tables list).  Q: list of bitmask rows (symmetric, zero diagonal)."""
    n = len(legs)
    if n > 28:
        raise ValueError("too many legs in one job")
    alpha = alpha or [0] * n
    beta = beta or [0] * n
    Q = Q or [0] * n
    dims = [L.n for L in legs]

    def stride_of(off):
        return int(off[1] - off[0]) if len(off) > 1 else 0
    if in_order is None:
        in_order = sorted(range(n), key=lambda a: -abs(stride_of(legs[a].in_off)))
    if out_order is None:
        out_order = sorted(range(n), key=lambda a: -abs(stride_of(legs[a].out_off)))
    groups, transpose = _choose_groups(dims, in_order, out_order)
    tabs = [_fuse(legs, g, alpha, beta, Q) for g in groups]
    sizes = [len(t) for t in tabs]
    nA = sizes[0]
    nB = sizes[1] if len(sizes) > 1 else 1
    ntiles = ((nA + TILE - 1) // TILE) * ((nB + TILE - 1) // TILE)
    for s in sizes[2:]:
        ntiles *= s
    return dict(in_base=int(in_base), out_base=int(out_base), sizes=sizes, const=int(const) & 1,
                conj=int(bool(conj)), transpose=int(transpose), ntiles=int(ntiles)), tabs


class PermutePlan:
    """Device-resident job + table arrays for one gtn_sign_permute launch."""

    def __init__(self, jobs):
        # jobs: list of (fields, tabs)
        self.njobs = len(jobs)
        self.max_tiles = 0
        self.elems = 0
        if self.njobs == 0:
            return
        for f, tabs in jobs:
            self.elems += int(np.prod([len(t) for t in tabs], dtype=np.int64))
        arr = (PermuteJob * self.njobs)()
        chunks, start = [], 0
        for k, (f, tabs) in enumerate(jobs):
            j = arr[k]
            j.in_base, j.out_base = f["in_base"], f["out_base"]
            for x in range(_cabi.GTN_MAX_SUPER):
                j.size[x] = 1
                j.table_start[x] = 0
            for x, t in enumerate(tabs):
                j.size[x] = len(t)
                j.table_start[x] = start
                start += len(t)
                chunks.append(t)
            j.nsuper = len(tabs)
            j.const_exp, j.conj, j.transpose, j.ntiles = f["const"], f["conj"], f["transpose"], f["ntiles"]
            self.max_tiles = max(self.max_tiles, f["ntiles"])
        self.jobs_dev = _to_dev_bytes(bytes(arr))
        self.entries_dev = _to_dev_bytes(np.concatenate(chunks).tobytes())

    def run(self, src, dst, scale=1.0):
        if self.njobs == 0:
            return
        code = dtype_code(src.dtype)
        s = complex(scale)
        # grid.y is limited to 65535 jobs; far above any block count we generate
        with prof_region("sign_permute", 1, 2 * self.elems * src.element_size()):
            check(lib.gtn_sign_permute(_ptr(src), _ptr(dst), code, _ptr(self.jobs_dev), _ptr(self.entries_dev),
                                       self.njobs, self.max_tiles, s.real, s.imag, _stream()), "gtn_sign_permute")


# ------------------------------------------------------------------------------------------------
#  block tensor
# ------------------------------------------------------------------------------------------------
_sigma_cache = {}


def sigma_bits(pi, n):
    """sigma exponent bit of block element (pi, s), s < n: bit 1 of popcount(canonical index)
    (reference sgn[blck][axis][sub_i] = param.sgn(param.encoder(i)), __init__.py:300-303).
    Returns a cached read-only array."""
    v = _sigma_cache.get((pi, n))
    if v is None:
        c = canonical_of_block(pi, np.arange(n, dtype=np.int64))
        v = ((_popcount_vec(c) >> 1) & 1).astype(np.uint32)
        v.setflags(write=False)
        if len(_sigma_cache) < 4096:
            _sigma_cache[(pi, n)] = v
    return v


def _register_dependent(src, bt):
    """src remembers (weakly) the unwritten permutations that read it: _ops.settle(src) writes them before src's
    storage is handed out for modification"""
    import weakref
    src.__dict__.setdefault("_deps", weakref.WeakSet()).add(bt)


class BT:
    """Parity-blocked Grassmann tensor on the device."""

    def __init__(self, stats, edims, odims, dtype, fmt="standard"):
        self.stats = tuple(stats)
        self.e = tuple(int(x) for x in edims)
        self.o = tuple(int(x) if s in FERMI else 0 for x, s in zip(odims, self.stats))
        self.dtype = dtype
        self.fmt = fmt
        self.faxes = tuple(a for a, s in enumerate(self.stats) if s in FERMI)
        self.off = {}        # pattern -> element offset (stored blocks only)
        self.zero = set()    # stored blocks known to be identically zero
        self._lazy = None    # a pending permutation of another tensor (_ops.LazyPermute): `off` describes the buffer
        self._buf = None     # it WILL have; consumers that pack anyway read the source and skip this pass

    # ---- storage
    @property
    def buf(self):
        b = self._buf
        if b is None and self._lazy is not None:
            lz, self._lazy = self._lazy, None
            self.__dict__.pop("_key", None)
            b = self._buf = lz.materialise()
        return b

    @buf.setter
    def buf(self, value):
        self._buf = value
        if self._lazy is not None:
            self._lazy = None
            self.__dict__.pop("_key", None)

    def pending(self):
        """the unwritten permutation this tensor stands for, or None"""
        return self._lazy if self._buf is None else None

    # ---- geometry
    @property
    def ndim(self):
        return len(self.stats)

    def patterns(self):
        return itertools.product((0, 1), repeat=len(self.faxes))

    def block_shape(self, pat):
        shp = list(self.e)
        for a, pi in zip(self.faxes, pat):
            shp[a] = self.e[a] if pi == 0 else self.o[a]
        return tuple(shp)

    def block_size(self, pat):
        return math.prod(self.block_shape(pat)) if self.ndim else 1

    def plan_blocks(self, pats=None):
        """lay the blocks `pats` out back to back (offsets only); returns the number of elements"""
        pats = list(self.patterns()) if pats is None else list(pats)
        total = 0
        for p in pats:
            self.off[p] = total
            total += self.block_size(p)
        return total

    def alloc(self, pats=None, zero=False):
        total = self.plan_blocks(pats)
        dev = require_cuda()
        self.buf = (torch.zeros if zero else torch.empty)(max(total, 1), dtype=self.dtype, device=dev)
        return self

    def live(self):
        """stored, not-known-zero, non-empty blocks"""
        return [p for p in self.off if p not in self.zero and self.block_size(p) > 0]

    def is_even(self):
        return all(sum(p) % 2 == 0 for p in self.live())

    def block_view(self, pat):
        shp = self.block_shape(pat)
        n = self.block_size(pat)
        if pat in self.off:
            return self.buf[self.off[pat]: self.off[pat] + n].view(shp)
        return torch.zeros(shp, dtype=self.dtype, device=self.buf.device if self.buf is not None else require_cuda())

    def key(self):
        """hashable description of the layout (cached: a BT's layout is not changed after it has
        been handed to an op)"""
        k = self.__dict__.get("_key")
        lz = self.pending()
        if k is None or k[0] != (len(self.off), len(self.zero), self.fmt, lz is None):
            k = ((len(self.off), len(self.zero), self.fmt, lz is None),
                 (self.stats, self.e, self.o, str(self.dtype), self.fmt,
                  tuple(sorted((p, o) for p, o in self.off.items())), tuple(sorted(self.zero)))
                 + (() if lz is None else (lz.sig,)))
            self._key = k
        return k[1]

    def stored_elems(self):
        """elements of the stored blocks (known without the buffer)"""
        return sum(self.block_size(p) for p in self.off)

    def layout_key(self):
        """key() of the written tensor (the same whether or not it is still an unwritten permutation)"""
        k = self.key()
        return k[:-1] if self.pending() is not None else k

    def leg(self, a):
        return (self.stats[a], self.e[a], self.o[a])

    def clone(self):
        r = BT(self.stats, self.e, self.o, self.dtype, self.fmt)
        r.off = dict(self.off)
        r.zero = set(self.zero)
        lz = self.pending()
        if lz is not None:
            r._lazy = lz                 # the copy of an unwritten permutation is the same unwritten permutation
            _register_dependent(lz.src, r)
        else:
            r.buf = self.buf.clone()
        return r

    def sumsq(self, pats=None):
        lz = self.pending()
        if lz is not None and pats is None:
            return lz.src.sumsq()        # a signed permutation has the norm of its source
        acc = torch.zeros(1, dtype=torch.float64, device=self.buf.device)
        code = dtype_code(self.dtype)
        covered = sum(self.block_size(p) for p in self.off)
        if pats is None and len(self.zero) == 0 and covered == self.buf.numel():
            # the stored blocks tile the buffer exactly: one launch
            check(lib.gtn_sumsq(_ptr(self.buf), self.buf.numel(), code, _ptr(acc), 0, _stream()), "gtn_sumsq")
            count()
        else:
            for p in (self.live() if pats is None else pats):
                if p in self.off and self.block_size(p) > 0:
                    v = self.buf[self.off[p]:]
                    check(lib.gtn_sumsq(_ptr(v), self.block_size(p), code, _ptr(acc), 0, _stream()), "gtn_sumsq")
                    count()
        return acc

    def norm(self):
        return math.sqrt(float(self.sumsq().item()))

    def sumabs(self, pats):
        """sum of |x| over the blocks `pats` (the reference's Grassmann-evenness tests are L1 means)"""
        acc = torch.zeros(1, dtype=torch.float64, device=self.buf.device)
        code = dtype_code(self.dtype)
        for p in pats:
            if p in self.off and self.block_size(p) > 0:
                v = self.buf[self.off[p]:]
                check(lib.gtn_sumabs(_ptr(v), self.block_size(p), code, _ptr(acc), 0, _stream()), "gtn_sumabs")
                count()
        return acc

    def scale_(self, s):
        s = complex(s)
        lz = self.pending()
        if lz is not None:               # written once, scaled on the way
            self._lazy = None
            self.__dict__.pop("_key", None)
            self._buf = lz.materialise(s.real if s.imag == 0.0 else s)
            return self
        check(lib.gtn_scale(_ptr(self.buf), self.buf.numel(), dtype_code(self.dtype), s.real, s.imag, _stream()),
              "gtn_scale")
        count()
        return self


def _row_strides(shape):
    st, acc = [0] * len(shape), 1
    for a in range(len(shape) - 1, -1, -1):
        st[a] = acc
        acc *= shape[a]
    return st


_plan_cache = {}
_graph_droppers = []        # callbacks that forget recorded CUDA graphs (gauge2d registers its step graphs)


def drop_graphs():
    """Forget every recorded CUDA graph (truncated-SVD schedules, whole-step graphs).  Recorded graphs bake raw
    pointers to the device launch tables of the plans they enqueued; those tables are owned by the plan caches
    (`_plan_cache`, `_ops._einsum_cache`), so a cache wipe must take the graphs with it -- otherwise a replay would
    read tables from freed (and possibly reused) memory."""
    for plan in list(globals().get("_trunc_plans", {}).values()):
        plan.graphs.clear()
        plan.graph_launches.clear()
        plan.warmed.clear()
    for fn in _graph_droppers:
        fn()


def _cached(key, builder):
    p = _plan_cache.get(key)
    if p is None:
        if len(_plan_cache) > 4096:
            _plan_cache.clear()
            drop_graphs()
        p = builder()
        _plan_cache[key] = p
    return p


# ---- dense <-> BT --------------------------------------------------------------------------
def _dense_leg_maps(d, stat, encoder):
    """for storage index i of a dense fermionic leg: (pi, s)."""
    i = np.arange(d, dtype=np.int64)
    if stat not in FERMI:
        return None, i
    j = i if encoder == "parity-preserving" else encoder_array(d)    # parity-preserving index
    return (j & 1), (j >> 1)


def bt_from_dense(data, stats, encoder="canonical", fmt="standard", check_even=True):
    """dense (canonical or parity-preserving) torch tensor -> BT; one sign-free launch.
    Replaces the element-wise Python loop of reference block.__init__ (__init__.py:311-337)."""
    shape = tuple(data.shape)
    e = [((d + 1) // 2 if s in FERMI else d) for d, s in zip(shape, stats)]
    o = [(d // 2 if s in FERMI else 0) for d, s in zip(shape, stats)]
    bt = BT(stats, e, o, data.dtype, fmt).alloc()
    if data.numel() == 0:
        return bt
    data = data.contiguous()

    def build():
        in_st = _row_strides(shape)
        legs = []
        # out offset of element = off[pattern] + sum s_a * stride_a(pattern); for power-of-two dims all
        # non-empty blocks have equal shape, so the offset is separable per leg
        pats = [p for p in bt.patterns() if bt.block_size(p) > 0]
        ref_shape = bt.block_shape(pats[0])
        if any(bt.block_shape(p) != ref_shape for p in pats):
            raise ValueError("Error[dense]: Some of the fermionic tensor shapes are not a power of two.")
        bstr = _row_strides(ref_shape)
        bsize = bt.block_size(pats[0])
        # weight of pi_a in the pattern index (patterns with an empty block never occur)
        weights, w = {}, 1
        for a in reversed(bt.faxes):
            if bt.o[a] > 0:
                weights[a] = w
                w *= 2
            else:
                weights[a] = 0
        # recompute offsets to match the separable formula
        for p in pats:
            bt.off[p] = sum(weights[a] * pi for a, pi in zip(bt.faxes, p)) * bsize
        for a, (d, s) in enumerate(zip(shape, stats)):
            pi, sidx = _dense_leg_maps(d, s, encoder)
            i = np.arange(d, dtype=np.int64)
            if pi is None:
                legs.append(LegTab(d, i * in_st[a], i * bstr[a]))
            else:
                legs.append(LegTab(d, i * in_st[a], pi * (weights[a] * bsize) + sidx * bstr[a]))
        order = list(range(len(shape)))
        return PermutePlan([build_job(legs, in_order=order, out_order=order)]), dict(bt.off)
    plan, off = _cached(("d2b", shape, tuple(stats), encoder, str(data.dtype)), build)
    bt.off = dict(off)
    for p in list(bt.off):
        if bt.block_size(p) == 0:
            del bt.off[p]
    plan.run(data.view(-1), bt.buf)
    if check_even and bt.faxes:
        odd = [p for p in bt.off if sum(p) % 2 == 1]
        if odd and float(bt.sumsq(odd).item()) == 0.0:
            bt.zero.update(odd)
    return bt


def _pad_pow2(n):
    return 1 if n <= 1 else 2 ** int(math.ceil(math.log2(n)))


def bt_dense_shape(bt):
    return tuple((_pad_pow2(bt.e[a] + bt.o[a]) if bt.stats[a] in FERMI else bt.e[a]) for a in range(bt.ndim))


def bt_to_dense(bt, encoder="canonical"):
    """BT -> dense torch tensor (fermionic dims padded to powers of two), one launch.
    Replaces reference todense (__init__.py:690-729)."""
    shape = bt_dense_shape(bt)
    live = bt.live()
    full = all(bt.e[a] + bt.o[a] == shape[a] for a in bt.faxes) and len(live) == 2 ** len(bt.faxes)
    dev = require_cuda()
    out = (torch.empty if full else torch.zeros)(shape, dtype=bt.dtype, device=dev)
    if out.numel() == 0 or not live:
        return out

    def build():
        ost = _row_strides(shape)
        jobs = []
        for p in live:
            bshape = bt.block_shape(p)
            bstr = _row_strides(bshape)
            pis = dict(zip(bt.faxes, p))
            legs = []
            for a in range(bt.ndim):
                s = np.arange(bshape[a], dtype=np.int64)
                if a in pis:
                    j = 2 * s + pis[a]
                    if encoder == "canonical":
                        j = canonical_of_block(pis[a], s)
                    legs.append(LegTab(bshape[a], s * bstr[a], j * ost[a]))
                else:
                    legs.append(LegTab(bshape[a], s * bstr[a], s * ost[a]))
            order = list(range(bt.ndim))
            jobs.append(build_job(legs, in_base=bt.off[p], in_order=order, out_order=order))
        return PermutePlan(jobs)
    plan = _cached(("b2d", bt.key(), encoder), build)
    plan.run(bt.buf, out.view(-1))
    return out


def dense_sign_permute(data, perm, alpha_axes=()):
    """out = transpose(data, perm) with (-1)^{popcount(i_a)} on the legs a in alpha_axes: the
    bosons-to-the-left step of join_legs / split_legs (reference __init__.py:3003-3015, :3153-3157)
    as ONE launch on a dense canonical tensor.  Here the parities are NOT block constants: the
    kernel evaluates the sign per element from the popcount tables of the legs."""
    shape = tuple(data.shape)
    n = len(shape)
    data = data.contiguous()
    out_shape = tuple(shape[a] for a in perm)
    out = torch.empty(out_shape, dtype=data.dtype, device=data.device)
    if out.numel() == 0:
        return out

    def build():
        in_st = _row_strides(shape)
        ost = _row_strides(out_shape)
        ost_of = {a: ost[k] for k, a in enumerate(perm)}
        legs, alpha = [], []
        for a in range(n):
            i = np.arange(shape[a], dtype=np.int64)
            par = (_popcount_vec(i) & 1).astype(np.uint32) if a in alpha_axes else None
            legs.append(LegTab(shape[a], i * in_st[a], i * ost_of[a], p=par))
            alpha.append(1 if a in alpha_axes else 0)
        return PermutePlan([build_job(legs, alpha=alpha, in_order=list(range(n)), out_order=list(perm))])
    plan = _cached(("dsp", shape, tuple(perm), tuple(sorted(alpha_axes)), str(data.dtype)), build)
    plan.run(data.view(-1), out.view(-1))
    return out


# ---- in-place style elementwise sign passes (format switch) -----------------------------------
def bt_switch_format(bt, sigma=None):
    """sigma on every conjugated (-1) leg: reference dense.switch_format (__init__.py:1011-1068) /
    block.switch_format (:537-571).  One launch, new buffer.
    sigma: {axis: (bits of the even block, bits of the odd block)} for legs whose sigma vector is not the standard
    one (legs made by join_legs_block carry their own sgn, reference :3588-3610)."""
    out = BT(bt.stats, bt.e, bt.o, bt.dtype, "matrix" if bt.fmt == "standard" else "standard")
    out.off = dict(bt.off)
    out.zero = set(bt.zero)
    out.buf = torch.empty_like(bt.buf)
    live = bt.live()
    if bt.zero:
        out.buf.zero_()

    def build():
        jobs = []
        for p in live:
            bshape = bt.block_shape(p)
            bstr = _row_strides(bshape)
            pis = dict(zip(bt.faxes, p))
            legs, beta = [], []
            for a in range(bt.ndim):
                if a in pis and bt.stats[a] == -1:
                    q = sigma[a][pis[a]][:bshape[a]] if (sigma and a in sigma) else sigma_bits(pis[a], bshape[a])
                    legs.append(lin_leg(bshape[a], bstr[a], bstr[a], q=q))
                    beta.append(1)
                else:
                    legs.append(lin_leg(bshape[a], bstr[a], bstr[a]))
                    beta.append(0)
            order = list(range(bt.ndim))
            jobs.append(build_job(legs, beta=beta, in_base=bt.off[p], out_base=bt.off[p], in_order=order, out_order=order))
        return PermutePlan(jobs)
    skey = None if not sigma else tuple((a, v[0].tobytes(), v[1].tobytes()) for a, v in sorted(sigma.items()))
    _cached(("fmt", bt.key(), skey), build).run(bt.buf, out.buf)
    return out


def bt_force_standard(bt):
    return bt if bt.fmt == "standard" else bt_switch_format(bt)


# ------------------------------------------------------------------------------------------------
#  group layouts (joined index spaces)
# ------------------------------------------------------------------------------------------------
_layout_cache = {}


def group_layout(legs, order="lex"):
    k = (tuple(legs), order)
    g = _layout_cache.get(k)
    if g is None:
        if len(_layout_cache) > 4096:
            _layout_cache.clear()
        g = GroupLayout(k[0], order)
        _layout_cache[k] = g
    return g


class GroupLayout:
    """Index space of a list of legs joined together.  Patterns (parities of the fermionic
    members) are ordered with even total parity first; inside a pattern the members form a
    row-major mixed-radix index.
    order='lex': patterns of one total parity in lexicographic order (internal packing: any fixed order
    serves a decomposition or a contraction).  order='ref': the order the reference's join_index walks the
    sub-blocks (__init__.py:3339-3355 through param.to_bin_parity_preserving, param.py:25-47): by
    (p_1, p_2, ...) read as a little-endian number, p_0 fixed by the total parity -- the user-visible layout
    of join_legs_block."""

    def __init__(self, legs, order="lex"):
        self.legs = list(legs)                      # (stat, e, o)
        self.fpos = [k for k, L in enumerate(self.legs) if L[0] in FERMI]
        pats = list(itertools.product((0, 1), repeat=len(self.fpos)))
        if order == "ref":
            pats.sort(key=lambda p: (sum(p) % 2, tuple(reversed(p[1:]))))
        else:
            pats.sort(key=lambda p: (sum(p) % 2, p))
        self.pats, self.offset, self.size, self.shape = [], {}, {}, {}
        acc = 0
        self.even_total = 0
        for p in pats:
            shp = [L[1] for L in self.legs]
            for k, pi in zip(self.fpos, p):
                shp[k] = self.legs[k][1] if pi == 0 else self.legs[k][2]
            sz = math.prod(shp) if shp else 1
            self.pats.append(p)
            self.offset[p] = acc
            self.size[p] = sz
            self.shape[p] = tuple(shp)
            acc += sz
            if sum(p) % 2 == 0:
                self.even_total = acc
        self.total = acc

    def sector(self, parity):
        """(start, length) of the rows with total parity `parity`"""
        return (0, self.even_total) if parity == 0 else (self.even_total, self.total - self.even_total)

    def strides(self, p):
        return _row_strides(self.shape[p])


# ------------------------------------------------------------------------------------------------
#  grouped GEMM
# ------------------------------------------------------------------------------------------------
SKINNY_MAX = int(__import__("os").environ.get("GTN_SKINNY_MAX", "128"))
# TMA-staged DMMA kernel (csrc/gtn_gemm_tma.cu) for launches with at least this many 128x64 (or 64x64) tiles
USE_TMA = bool(int(__import__("os").environ.get("GTN_TMA", "1")))
TMA_MIN_TILES = int(__import__("os").environ.get("GTN_TMA_MIN_TILES", "37"))
TMA_MAX_GROUPS = 32         # GTN_TMA_MAX_GROUPS of include/gtn_b200.h
GEMM_FAMILY = {0: "gemm_ldgsts_64x64", 1: "gemm_skinny_32x32", 3: "gemm_skinny_32x32", 4: "gemm_tma_64x64",
               12: "gemm_tma_128x64"}


SPLIT_K = bool(int(__import__("os").environ.get("GTN_SPLIT_K", "1")))


class GemmPlan:
    """Device-resident group list of one grouped GEMM launch and the tile configuration it runs in:
       12 / 4  TMA-staged 128x64 / 64x64 tiles (large products: every side >= 64, enough tiles to fill the GPU,
               16-byte aligned operands, <= 32 groups), tile grid rasterised in bands of row tiles;
       1       32x32 tiles with a deep K step (skinny products and Gram matrices of the subspace iteration);
       0       64x64 cp.async tiles (everything else: ragged block sectors, misaligned float64 views)."""

    def __init__(self, groups, dtype, config=None, splitk=True):
        self.n = len(groups)
        self.tiles = 0
        self.flops = 0
        self.bytes = 0
        self.slices = 1
        if self.n == 0:
            return
        cplx = dtype == torch.complex128
        if splitk and config is None and SPLIT_K:
            groups = self._split_k(groups, dtype)
            self.n = len(groups)
        for g in groups:
            b = g.get("batch", 1)
            self.flops += (8 if cplx else 2) * b * g["m"] * g["n"] * g["k"]
            self.bytes += (16 if cplx else 8) * b * (g["m"] * g["k"] + g["k"] * g["n"] + g["m"] * g["n"])
        self.desc = "%d groups, first %dx%dx%d, slices %d" % (self.n, groups[0]["m"], groups[0]["n"], groups[0]["k"],
                                                               self.slices)
        arr = (GemmGroup * self.n)()
        for k, g in enumerate(groups):
            a = arr[k]
            a.a_off, a.b_off, a.c_off = g["a_off"], g["b_off"], g["c_off"]
            a.lda, a.ldb, a.ldc = g["lda"], g["ldb"], g["ldc"]
            a.batch_stride_a = g.get("bsa", 0)
            a.batch_stride_b = g.get("bsb", 0)
            a.batch_stride_c = g.get("bsc", 0)
            a.m, a.n, a.k, a.batch = g["m"], g["n"], g["k"], g.get("batch", 1)
            a.alpha, a.beta = g.get("alpha", 1.0), g.get("beta", 0.0)
            a.flags = g.get("flags", 0)
        code = dtype_code(dtype)
        if config is None:
            config = self._choose(groups, arr, code)
        if any(g.get("flags", 0) & 1 for g in groups):
            # conj-transposed B operand: a whole-launch property, only built for the 32x32 configuration
            assert all(g.get("flags", 0) & 1 for g in groups)
            config = 3
        self.config = config
        if config & 4:
            # band height of the rasterised tile walk: the CTAs resident at one time (148 x 128x64 tiles, or 296 x
            # 64x64) cover a square-ish region of C
            for a in arr:
                a.reserved = 8 if config & 8 else 16
        self.family = GEMM_FAMILY.get(config, "grouped_gemm")
        self.tiles = int(lib.gtn_gemm_plan_host(arr, self.n, code, self.config))
        self.host = arr                      # the TMA path encodes its tensor maps from the host copy
        self.dev = _to_dev_bytes(bytes(arr))

    def _split_k(self, groups, dtype):
        """Gram-type launches -- a handful of output tiles over a very long contracted range (the environment matrices
        of hotrg3dz: 64 x 64 outputs, K = 131072 per sector: 4 CTAs at work for a millisecond) -- are cut into
        `slices` ranges of K that write consecutive copies of the output; gtn_sum_slices adds them in a fixed order.
        Only for compact outputs (ldc = n, the groups tile one contiguous span) with beta = 0 and no batch."""
        if any(g.get("beta", 0.0) != 0.0 or g.get("batch", 1) != 1 or g.get("flags", 0) or g["ldc"] != g["n"]
               for g in groups):
            return groups
        tiles = sum(-(-g["m"] // 32) * -(-g["n"] // 32) for g in groups)
        kmax = max(g["k"] for g in groups)
        if tiles * 4 > 148 or kmax < 4096:
            return groups
        lo = min(g["c_off"] for g in groups)
        span = sum(g["m"] * g["n"] for g in groups)
        if max(g["c_off"] + g["m"] * g["n"] for g in groups) - lo != span:
            return groups
        if dtype != torch.complex128 and span % 2:
            return groups
        slices = int(min(64, max(2, 296 // tiles), kmax // 1024))
        if slices < 2:
            return groups
        out = []
        for s_ in range(slices):
            for g in groups:
                kc = -(-g["k"] // slices)
                k0 = min(s_ * kc, g["k"])
                kk = min(kc, g["k"] - k0)
                h = dict(g)
                h["a_off"], h["b_off"] = g["a_off"] + k0, g["b_off"] + k0 * g["ldb"]
                h["c_off"] = s_ * span + (g["c_off"] - lo)
                if kk <= 0:
                    h["k"], h["alpha"] = 1, 0.0
                    h["a_off"], h["b_off"] = g["a_off"], g["b_off"]
                else:
                    h["k"] = kk
                out.append(h)
        self.slices, self.span, self.c_lo = slices, span, lo
        return out

    # measured on the B200 (scripts/gemm_bench.py, profiles/r2_gemm_bench.json): rate of a configuration on a large
    # problem without a partial last wave, relative to cuBLAS ZGEMM; CTAs resident per GPU; tile shape
    _CFG = {12: (0.975, 148, 128, 64), 4: (0.975, 296, 64, 64), 1: (0.91, 296, 32, 32), 0: (0.95, 296, 64, 64)}

    def _choose(self, groups, arr, code):
        """the configuration with the best estimated rate: (rate on full waves) x (useful part of the padded tiles)
        x (filled part of the last wave).  Large square products take the TMA tiles; the l x q x p panels of the
        subspace iteration at chi = 128 (512 tiles of 128x64: 3.5 waves) run faster on the fine 32x32 tiles."""
        cands = [1, 0]
        if USE_TMA and self.n <= TMA_MAX_GROUPS and min(min(g["m"], g["n"]) for g in groups) >= 64 \
                and min(g["k"] for g in groups) >= 16 and lib.gtn_gemm_tma_check(arr, self.n, code):
            cands.insert(0, 12 if max(g["m"] for g in groups) >= 128 else 4)
        useful = sum(g["m"] * g["n"] * g.get("batch", 1) for g in groups)
        best, best_eff = 0, -1.0
        for cfg in cands:
            base, slots, bm, bn = self._CFG[cfg]
            tiles = int(lib.gtn_gemm_plan_host(arr, self.n, code, cfg))
            if tiles <= 0:
                continue
            waves = -(-tiles // slots)
            eff = base * (useful / (tiles * bm * bn)) * (tiles / (waves * slots))
            if (cfg & 4) and tiles < TMA_MIN_TILES:
                continue
            if eff > best_eff * 1.005:
                best, best_eff = cfg, eff
        return best

    def run(self, A, B, Cm):
        if self.n == 0 or self.tiles == 0:
            return
        if self.slices > 1:
            tmp = torch.empty(self.slices * self.span, dtype=Cm.dtype, device=Cm.device)
            self._launch(A, B, tmp)
            with prof_region("sum_slices", 1, (self.slices + 1) * self.span * tmp.element_size()):
                dst = C.c_void_p(Cm.data_ptr() + self.c_lo * Cm.element_size())
                check(lib.gtn_sum_slices(_ptr(tmp), dst, self.span, self.slices, dtype_code(Cm.dtype), _stream()),
                      "gtn_sum_slices")
            return
        self._launch(A, B, Cm)

    def _launch(self, A, B, Cm):
        with prof_region(self.family, 1, self.bytes, self.flops, self.desc):
            if self.config & 4:
                check(lib.gtn_grouped_gemm_tma(_ptr(A), _ptr(B), _ptr(Cm), dtype_code(A.dtype), self.host, _ptr(self.dev),
                                               self.n, self.tiles, self.config, _stream()), "gtn_grouped_gemm_tma")
            else:
                check(lib.gtn_grouped_gemm(_ptr(A), _ptr(B), _ptr(Cm), dtype_code(A.dtype), _ptr(self.dev), self.n,
                                           self.tiles, self.config, _stream()), "gtn_grouped_gemm")


def gemm(A, B, m, n, k, lda=None, ldb=None, ldc=None, out=None, config=None):
    """plain C[m,n] = A[m,k] B[k,n] on the DMMA kernel (row-major)."""
    if out is None:
        out = torch.empty(m * n, dtype=A.dtype, device=A.device)
    plan = _cached(("gemm1", m, n, k, lda, ldb, ldc, str(A.dtype), config), lambda: GemmPlan(
        [dict(a_off=0, b_off=0, c_off=0, lda=lda or k, ldb=ldb or n, ldc=ldc or n, m=m, n=n, k=k)], A.dtype, config))
    plan.run(A, B, out)
    return out


# ------------------------------------------------------------------------------------------------
#  batched Jacobi SVD
# ------------------------------------------------------------------------------------------------
JACOBI_TOL = 4e-15
JACOBI_MAX_SWEEPS = 60
# projected matrix of the truncated SVD: skip the confirming sweep once a sweep only rotated pairs with a normalised
# inner product below this (quadratic convergence; the residual certificate checks the result anyway).  0 = off
JACOBI_EARLY_STOP = float(__import__("os").environ.get("GTN_JACOBI_EARLY_STOP", "1e-10"))
PERSISTENT_MAX_ROWS = 160   # batches up to this many rows try the one-launch cooperative Jacobi kernel


def batched_svd(mats):
    """mats: list of 2-D device tensors (row-major, any shapes).  Returns list of (U, s, Vh) with
    U (m x r), s (r, host numpy, descending), Vh (r x n), r = min(m, n); all through the Jacobi
    kernels in csrc/gtn_svd.cu.  Replaces np.linalg.svd of reference SortedSVD (__init__.py:3932)."""
    if not mats:
        return []
    dev = mats[0].device
    dt = mats[0].dtype
    code = dtype_code(dt)
    probs, works, transposed = [], [], []
    woff = zoff = soff = 0
    for Mx in mats:
        m, n = Mx.shape
        tr = m > n
        Wm = (Mx.transpose(0, 1) if tr else Mx).contiguous()
        p, q = Wm.shape
        probs.append((woff, zoff, p, q, soff))
        works.append(Wm.reshape(-1))
        transposed.append(tr)
        woff += p * q
        zoff += p * p
        soff += p
    nprob = len(mats)
    maxp = max(pr[2] for pr in probs)
    maxq = max(pr[3] for pr in probs)
    W = torch.cat(works) if nprob > 1 else works[0].clone()
    if W.numel() == 0:
        W = torch.zeros(1, dtype=dt, device=dev)
    Z = torch.empty(max(zoff, 1), dtype=dt, device=dev)
    parr = (SvdProblem * nprob)()
    oarr = (SvdOut * nprob)()
    for k, (wo, zo, p, q, so) in enumerate(probs):
        parr[k].w_off, parr[k].z_off, parr[k].p, parr[k].q = wo, zo, p, q
        oarr[k].s_off, oarr[k].u_off = so, zo
    pdev = _to_dev_bytes(bytes(parr))
    odev = _to_dev_bytes(bytes(oarr))
    st = _stream()
    rn2 = torch.empty(max(soff, 1), dtype=torch.float64, device=dev)
    fro2 = torch.empty(2 * nprob, dtype=torch.float64, device=dev)      # double buffered by round parity
    rn_off = torch.tensor([pr[4] for pr in probs], dtype=torch.int64).to(dev, non_blocking=True)
    check(lib.gtn_jacobi_init(_ptr(W), _ptr(Z), code, _ptr(pdev), nprob, maxp, _ptr(rn2), _ptr(fro2), _ptr(rn_off),
                              st), "gtn_jacobi_init")
    count()
    offd = torch.zeros(2 * nprob, dtype=torch.float64, device=dev)
    sweeps = 0
    done = False
    if 2 <= maxp <= PERSISTENT_MAX_ROWS:
        # small problems: the whole sweep loop in one cooperative launch
        sw = torch.zeros(4, dtype=torch.int32, device=dev)
        P = (maxp + 1) & ~1
        esz = W.element_size()
        rb = sum(2 * esz * (pr[2] * pr[3] + pr[2] * pr[2]) for pr in probs)
        with prof_region("jacobi_persistent", 1, 0) as pr_:
            rc = lib.gtn_jacobi_persistent(_ptr(W), _ptr(Z), code, _ptr(pdev), nprob, maxp, JACOBI_TOL, _ptr(offd),
                                           _ptr(rn2), _ptr(fro2), _ptr(rn_off), JACOBI_MAX_SWEEPS, _ptr(sw), st, 0.0)
            swh = sw.cpu().tolist()[:2] if rc == 0 else [0, 0]
            # algorithmic bytes: every row of W and Z read + written once per round
            pr_.set_bytes(rb * (P - 1) * max(swh[0], 1))
        if rc == 0:
            sweeps = swh[0]
            if not swh[1]:
                raise _cabi.GtnError("Jacobi SVD did not converge in %d sweeps" % sweeps)
            done = True
        elif rc != -2:
            check(rc, "gtn_jacobi_persistent")
    if maxp >= 2 and not done:
        while True:
            P = (maxp + 1) & ~1
            esz = W.element_size()
            rb = sum(2 * esz * (pr[2] * pr[3] + pr[2] * pr[2]) for pr in probs)   # every row of W and Z read+written
            with prof_region("jacobi_round", P - 1, rb * (P - 1)):
                check(lib.gtn_jacobi_sweep(_ptr(W), _ptr(Z), code, _ptr(pdev), nprob, maxp, maxq, JACOBI_TOL,
                                           _ptr(offd), _ptr(rn2), _ptr(fro2), _ptr(rn_off), st), "gtn_jacobi_sweep")
            sweeps += 1
            if float(offd[:nprob].max().item()) <= JACOBI_TOL ** 2:     # squared overlaps of rotated pairs
                break
            if sweeps >= JACOBI_MAX_SWEEPS:
                raise _cabi.GtnError("Jacobi SVD did not converge in %d sweeps" % sweeps)
    U = torch.empty_like(Z)
    Vh = torch.empty_like(W)
    s = torch.empty(max(soff, 1), dtype=torch.float64, device=dev)
    order = torch.empty(max(soff, 1), dtype=torch.int32, device=dev)
    scratch = torch.empty(max(soff, 1), dtype=torch.float64, device=dev)
    check(lib.gtn_jacobi_finish(_ptr(W), _ptr(Z), _ptr(U), _ptr(Vh), _ptr(s), code, _ptr(pdev), _ptr(odev),
                                _ptr(order), _ptr(scratch), nprob, maxp, maxq, st), "gtn_jacobi_finish")
    count(2)
    s_host = s.cpu().numpy()
    outs = []
    for (wo, zo, p, q, so), tr in zip(probs, transposed):
        Uk = U[zo: zo + p * p].view(p, p)
        Vk = Vh[wo: wo + p * q].view(p, q)
        sk = s_host[so: so + p].copy()
        if tr:
            # W = A^T = Uk diag(s) Vk  =>  A = Vk^T diag(s) Uk^T
            outs.append((Vk.transpose(0, 1), sk, Uk.transpose(0, 1)))
        else:
            outs.append((Uk, sk, Vk))
    batched_svd.last_sweeps = sweeps
    return outs


batched_svd.last_sweeps = 0


# ------------------------------------------------------------------------------------------------
#  truncated SVD: randomized subspace iteration (GEMM-bound) + small Jacobi + residual certificate
# ------------------------------------------------------------------------------------------------
TRUNC_TOL = 1e-11          # certificate: max_i ||W^H u_i - s_i v_i|| <= TRUNC_TOL * s_0
# singular values at or below RANK_NOISE * s_0 are within 10x of the null-space noise of the Jacobi kernel and of the
# projected matrix (measured up to 1.2e-14 s_0): whether they pass the reference's rank rule s_i / s_0 > 1e-14 is decided
# by the rank certificate (deflated_norm_bound), not by their computed values
RANK_NOISE = 1e-13
TRUNC_MAX_ITERS = 20
TRUNC_LMAX = 320           # widest subspace (rows of the projected matrix); gtn_chol_whiten handles <= 512
_rand_cache = {}


def _randn(n, dtype, dev):
    key = (n, str(dtype), str(dev))
    g = _rand_cache.get(key)
    if g is None:
        gen = torch.Generator(device="cpu")
        gen.manual_seed(20240607 + n)
        if dtype == torch.complex128:
            r = torch.randn(n, 2, generator=gen, dtype=torch.float64)
            g = torch.view_as_complex(r).to(dev)
        else:
            g = torch.randn(n, generator=gen, dtype=torch.float64).to(dev)
        if len(_rand_cache) < 64:
            _rand_cache[key] = g
    return g


# ------------------------------------------------------------------------------------------------
#  rank certificate: is everything outside the dominant left singular subspace below the rank rule's threshold?
# ------------------------------------------------------------------------------------------------
def _gemm_t(A, lda, B, ldb, Cm, ldc, m, n, k, alpha=1.0, beta=0.0):
    """Cm (m x n, row stride ldc) = alpha * A (m x k, lda) B (k x n, ldb) + beta * Cm on three separate buffers"""
    key = ("gemm_t", m, n, k, lda, ldb, ldc, alpha, beta, str(A.dtype))
    _cached(key, lambda: GemmPlan([dict(a_off=0, b_off=0, c_off=0, lda=lda, ldb=ldb, ldc=ldc, m=m, n=n, k=k,
                                        alpha=alpha, beta=beta)], A.dtype)).run(A, B, Cm)


def _ctranspose_t(src, r, c, ld, dst):
    """dst (c x r, contiguous) = src[:r, :c]^H (row stride ld), one sign+permute launch (conj + transpose)"""
    def build():
        legs = [lin_leg(r, ld, 1), lin_leg(c, 1, r)]
        return PermutePlan([build_job(legs, conj=(src.dtype == torch.complex128), in_order=[0, 1], out_order=[1, 0])])
    _cached(("ctrans_t", r, c, ld, str(src.dtype)), build).run(src.view(-1), dst.view(-1))


def _sumsq_t(t):
    acc = torch.zeros(1, dtype=torch.float64, device=t.device)
    check(lib.gtn_sumsq(_ptr(t), t.numel(), dtype_code(t.dtype), _ptr(acc), 0, _stream()), "gtn_sumsq")
    count()
    return acc


def deflated_norm_bound(W, UhD, U, ldu, nD, thr, allreduce=None):
    """Upper bound of || (I - U_D U_D^H) W ||_2 for W (p x q, contiguous), UhD (nD x p, contiguous rows u_i^H) and
    U (p x *, row stride ldu, first nD columns u_i): the part of W outside its nD dominant left singular directions.

    Why: the reference keeps the singular values with s_i / (s_0 + 1e-14) > 1e-14 of a LAPACK SVD
    (__init__.py:3939-3941), and the Z2 gauge tensors are exactly rank deficient early in a chain (16 of 512 per
    sector in the second TRG step at chi = 64).  LAPACK returns the null directions at 1e-16 ... 5e-16 s_0; the
    one-sided Jacobi kernel and the projected matrix of the subspace iteration return them at 2e-15 ... 1.2e-14 s_0
    (thousands of rotations / a p-term dot product per element), so their count of "non-zero" values is decided
    by rounding.  The deflated matrix N = W - U_D (U_D^H W), computed afresh from W in two projection passes, only
    carries W's own null part plus eps ||W||_F / sqrt(p)-sized rounding: if even its norm is below the threshold,
    no direction beyond the nD dominant ones can pass the rank rule in ANY implementation.
    Bounds: ||N||_2 <= ||N||_F, and when that is not conclusive (the noise of thousands of null directions adds up
    in the Frobenius norm) ||N||_2 <= ||N N^H||_F^(1/2) (Schatten-4 norm: one more GEMM on the smaller side).
    Column-sharded W (allreduce given): N_r is local, ||N||_F^2 = sum_r ||N_r||_F^2 and
    ||N||_2^2 <= sum_r ||N_r||_2^2 <= sum_r ||N_r^H N_r||_F."""
    p, q = W.shape
    dt, dev = W.dtype, W.device
    N = W.clone().view(-1)
    X = torch.empty(max(nD * q, 1), dtype=dt, device=dev)
    for _ in range(2 if nD > 0 else 0):                 # "twice is enough": U_D is orthonormal only to ~1e-14
        _gemm_t(UhD, p, N, q, X, q, nD, q, p)
        _gemm_t(U, ldu, X, q, N, q, p, q, nD, alpha=-1.0, beta=1.0)
    f2 = _sumsq_t(N)
    if allreduce is not None:
        allreduce(f2)
    f = math.sqrt(max(float(f2.item()), 0.0))
    if f <= thr or min(p, q) < 2:
        return f
    if f > thr * math.sqrt(min(p, q) * (1 if allreduce is None else 64)):
        # ||N||_2 >= ||N||_F / sqrt(rank): no bound can come below thr -- skip the min(p, q)^2 max(p, q) GEMM
        # (8192^3: 122 ms at chi = 128).  (Sharded: the number of ranks is not known here; 64 covers a box.)
        return f
    Nh = torch.empty(p * q, dtype=dt, device=dev)
    _ctranspose_t(N, p, q, q, Nh)                        # q x p
    if p <= q:
        G = torch.empty(p * p, dtype=dt, device=dev)
        _gemm_t(N, q, Nh, p, G, p, p, p, q)
    else:
        G = torch.empty(q * q, dtype=dt, device=dev)
        _gemm_t(Nh, p, N, q, G, q, q, q, p)
    t = torch.sqrt(_sumsq_t(G))
    if allreduce is not None:
        allreduce(t)
    return min(f, math.sqrt(max(float(t.item()), 0.0)))


RANK_CHECK_STATS = {"calls": 0, "certified": 0}


def refine_null_band(M, usv):
    """Full-SVD results (batched_svd) whose smallest values that pass the reference's rank rule lie inside the noise
    band (1e-14, 1e-13] s_0 of the Jacobi kernel: when the deflated matrix certifies that nothing beyond the values
    above the band can pass the rule (deflated_norm_bound), the values in the band are lowered to that bound -- the
    rank then equals LAPACK's on the same matrix.  Not conclusive: the result is left as the kernel delivered it."""
    U, sv, Vh = usv
    if len(sv) == 0 or not sv[0] > 0:
        return usv
    s0 = float(sv[0])
    nnz = int(np.sum(sv / (s0 + 1e-14) > 1e-14))
    if nnz == 0 or sv[nnz - 1] > RANK_NOISE * s0:
        return usv
    nD = int(np.sum(sv > RANK_NOISE * s0))
    thr = 1e-14 * (s0 + 1e-14)
    RANK_CHECK_STATS["calls"] += 1
    Mc = M.contiguous()
    p, q = Mc.shape
    Uc = U.contiguous()
    dt, dev = Mc.dtype, Mc.device
    UhD = torch.empty(max(nD * p, 1), dtype=dt, device=dev)
    _ctranspose_t(Uc, p, nD, Uc.shape[1], UhD)
    # The Jacobi kernel's dominant left vectors carry its backward error (thousands of rotations: the subspace is off
    # by ~1e-13, measured: the deflated matrix then has norm 2e-13 s_0 and nothing can be certified).  One step of
    # subspace iteration on the dominant subspace, Yh = (Uh_D M) M^H re-orthonormalised (the small nD x p panel through
    # the Jacobi kernels themselves), removes the leak to second order before the deflation.
    Z = torch.empty(nD * q, dtype=dt, device=dev)
    _gemm_t(UhD, p, Mc, q, Z, q, nD, q, p)
    Mh = torch.empty(p * q, dtype=dt, device=dev)
    _ctranspose_t(Mc, p, q, q, Mh)
    Yh = torch.empty(nD * p, dtype=dt, device=dev)
    _gemm_t(Z, q, Mh, p, Yh, p, nD, p, q)
    del Z, Mh
    (_, sy, Qh), = batched_svd([Yh.view(nD, p)])
    if len(sy) < nD or not sy[nD - 1] > 0:
        return usv
    Qh = Qh.contiguous()
    Qm = torch.empty(p * nD, dtype=dt, device=dev)
    _ctranspose_t(Qh, nD, p, p, Qm)                              # p x nD: the refined dominant left vectors
    bound = deflated_norm_bound(Mc, Qh, Qm, nD, nD, thr)
    if bound <= thr:
        RANK_CHECK_STATS["certified"] += 1
        sv = sv.copy()
        sv[nD:] = np.minimum(sv[nD:], bound)
        return (U, sv, Vh)
    return usv


class _WS:
    """carves matrices out of one device buffer so that grouped launches share a base pointer"""

    def __init__(self, dtype, dev):
        self.dtype, self.dev, self.items, self.total = dtype, dev, [], 0

    def add(self, rows, cols):
        off = self.total
        self.total += rows * cols
        self.items.append((off, rows, cols))
        return len(self.items) - 1

    def alloc(self):
        self.buf = torch.empty(max(self.total, 1), dtype=self.dtype, device=self.dev)

    def view(self, h):
        off, r, c = self.items[h]
        return self.buf[off: off + r * c].view(r, c)

    def off(self, h):
        return self.items[h][0]


def _ws_gemm(ws, triples, alpha=1.0, beta=0.0):
    """C_h = alpha * A_h B_h + beta * C_h for handles (a, b, c) inside ws, ONE grouped launch."""
    groups = []
    for a, b, c in triples:
        _, m, k = ws.items[a]
        _, k2, n = ws.items[b]
        assert k == k2 and ws.items[c][1:] == (m, n)
        groups.append(dict(a_off=ws.off(a), b_off=ws.off(b), c_off=ws.off(c), lda=k, ldb=n, ldc=n, m=m, n=n, k=k,
                           alpha=alpha, beta=beta))
    key = ("wsgemm", str(ws.dtype), tuple(tuple(sorted(g.items())) for g in groups))
    _cached(key, lambda: GemmPlan(groups, ws.dtype)).run(ws.buf, ws.buf, ws.buf)


def _ws_ctranspose(ws, pairs):
    """dst_h = src_h^H for handle pairs inside ws, one sign-permute launch (conj + transpose)."""
    def build():
        jobs = []
        for src, dst in pairs:
            _, r, c = ws.items[src]
            legs = [lin_leg(r, c, 1), lin_leg(c, 1, r)]
            jobs.append(build_job(legs, conj=(ws.dtype == torch.complex128), in_base=ws.off(src), out_base=ws.off(dst),
                                  in_order=[0, 1], out_order=[1, 0]))
        return PermutePlan(jobs)
    key = ("wsct", str(ws.dtype), tuple((ws.items[s], ws.items[d]) for s, d in pairs))
    _cached(key, build).run(ws.buf, ws.buf)


DEBUG_TRUNC = bool(int(__import__("os").environ.get("GTN_DEBUG_TRUNC", "0")))
WHITEN = "chol"            # "chol": pivoted Cholesky kernel (default); "eigh": Jacobi eigen-solver kernel
USE_GRAPHS = bool(int(__import__("os").environ.get("GTN_GRAPHS", "1")))
GEMM_CONJ_TRANS = bool(int(__import__("os").environ.get("GTN_GEMM_CONJ_TRANS", "0")))
PREROTATE = bool(int(__import__("os").environ.get("GTN_PREROTATE", "1")))
# in-kernel Jacobi tolerance of the Gram pre-rotation: the Gram matrix only determines the rows to ~1e-8
# relative for the small ones, the global Jacobi kernel polishes the rest
ROTATE_TOL = float(__import__("os").environ.get("GTN_ROTATE_TOL", "1e-12"))
TRUNC_PLAN_CACHE_BYTES = 1 << 30     # plans (workspace + CUDA graphs) are kept for workspaces up to this size
_trunc_iters_hint = {}
_os_env = __import__("os").environ.get("GTN_TRUNC_OVERSAMPLE")
TRUNC_OVERSAMPLE = int(_os_env) if _os_env else None


def subspace_rows(k):
    """Rows l of the iterated subspace for k wanted triplets: about 2k, snapped to what the kernels tile
    without padding -- 32 or 48 rows for k <= 16 (32-row GEMM tiles, register-tile Cholesky up to 48), a
    multiple of 64 beyond (64-row GEMM tiles).  Measured on the Z2 TRG chain: chi=32 (k=16) 48 rows 4.1 ms/step
    against 4.7 at 40 rows (fewer iterations, same padded GEMM cost); chi=64 (k=32) 64 rows 19.5 ms against
    22.3 at 72 rows (one GEMM row tile instead of two)."""
    if TRUNC_OVERSAMPLE is not None:
        return 2 * k + TRUNC_OVERSAMPLE
    if k <= 16:
        return 32 if 2 * k + 16 <= 32 else 48
    return 64 * (-(-2 * k // 64))


_trunc_fail = {}
_trunc_rate = {}
_trunc_robust = {}         # (batch shape, call site) -> True: the Gram-whitened iteration dropped directions here
ROBUST_SHIFT = 1e-9        # relative diagonal shift (x trace) of the shifted Cholesky QR passes of the robust mode
ROBUST_MIN_DIM = 1024      # sectors at least this large retry with Jacobi orthonormalisation before the full SVD
SVD_SITE = [None]          # call site of the decomposition being run (set by _ops.decompose_many)
FORCE_VERIFY_FAIL = [None]  # test hook: callable(SpeculativeSVD) -> True makes verify() report a failed certificate
CAPTURING_STEP = [False]   # a caller is capturing a whole coarse-graining step: enqueue the steady-state schedule inline


class NotCapturable(RuntimeError):
    """raised inside a whole-step capture when a decomposition is not in its replayable steady state"""
_trunc_plans = {}


class _TruncPlan:
    """Static workspace, device metadata and launch sequence of one truncated-SVD batch shape.
    Every step (start / iterate / check) only enqueues launches on the current stream; the one host
    read-back per check happens in read().  Because nothing in a step depends on host data, the
    steady-state schedule 'start, n iterations, check' is captured once per n as a CUDA graph and
    replayed -- ~75 launches and four host round trips of a chi=32 TRG step become one graph launch and
    one read-back."""

    def __init__(self, P_, Q_, ks, L_, dt, dev):
        self.P_, self.Q_, self.ks, self.L_, self.dt, self.dev = P_, Q_, ks, L_, dt, dev
        nb = self.nb = len(P_)
        ws = self.ws = _WS(dt, dev)
        add = lambda rows, cols: [ws.add(r, c) for r, c in zip(rows, cols)]
        self.hW, self.hWh = add(P_, Q_), add(Q_, P_)
        self.hG, self.hYh, self.hQh = add(L_, Q_), add(L_, P_), add(L_, P_)
        self.hZh, self.hPh = add(L_, Q_), add(L_, Q_)
        self.hB, self.hVk = add(L_, Q_), add(L_, Q_)          # same sizes, same order: constant offset delta
        self.hZ, self.hUb, self.hUbH = add(L_, L_), add(L_, L_), add(L_, L_)
        self.hUh, self.hU, self.hXh = add(L_, P_), add(P_, L_), add(L_, P_)
        # split-K Gram matrices: NS partial l x l slices per problem, summed by the whitening kernels
        self.NS = 4 if (WHITEN == "chol" and min(P_ + Q_) >= 256) else 1
        self.hD, self.hT1, self.hT2 = add(L_, L_), add([self.NS * l for l in L_], L_), add(L_, L_)
        self.hCp, self.hCq = add(P_, L_), add(Q_, L_)         # conj-transposed iterates
        self.hSp, self.hSq = add(L_, P_), add(L_, Q_)
        ws.alloc()
        self.nbytes = ws.buf.numel() * ws.buf.element_size()
        for b in range(nb):
            ws.view(self.hG[b]).view(-1).copy_(_randn(L_[b] * Q_[b], dt, dev))
            ws.view(self.hD[b]).zero_()
        sumL = self.sumL = sum(L_)
        self.soff = list(np.cumsum([0] + L_[:-1]))
        i64 = lambda v: torch.tensor(list(v), dtype=torch.int64).to(dev)
        # whitening
        self.g_off, self.t_off = i64(ws.off(h) for h in self.hT1), i64(ws.off(h) for h in self.hT2)
        self.e_off = i64(self.soff)
        self.n_dev = torch.tensor(L_, dtype=torch.int32).to(dev)
        self.kept = [torch.zeros(nb, dtype=torch.int32, device=dev) for _ in range(2)]
        self.evals = torch.empty(sumL, dtype=torch.float64, device=dev)
        se = int(lib.gtn_chol_whiten_scratch_elems(max(L_)))
        self.chol_scratch = torch.empty(se * nb, dtype=torch.complex128, device=dev) if se else None
        # small Jacobi SVD of B, in place inside the workspace
        parr, oarr = (SvdProblem * nb)(), (SvdOut * nb)()
        for b in range(nb):
            parr[b].w_off, parr[b].z_off, parr[b].p, parr[b].q = ws.off(self.hB[b]), ws.off(self.hZ[b]), L_[b], Q_[b]
            oarr[b].s_off, oarr[b].u_off = self.soff[b], ws.off(self.hUb[b])
        self.pdev, self.odev = _to_dev_bytes(bytes(parr)), _to_dev_bytes(bytes(oarr))
        self.vh_delta = ws.off(self.hVk[0]) - ws.off(self.hB[0])
        assert all(ws.off(v) - ws.off(h) == self.vh_delta for v, h in zip(self.hVk, self.hB))
        self.rn2 = torch.empty(sumL, dtype=torch.float64, device=dev)
        self.fro2 = torch.empty(2 * nb, dtype=torch.float64, device=dev)
        self.rn_off = i64(self.soff)
        self.offd = torch.zeros(2 * nb, dtype=torch.float64, device=dev)
        self.sw = torch.zeros(4, dtype=torch.int32, device=dev)
        self.s_dev = torch.empty(sumL, dtype=torch.float64, device=dev)
        self.order = torch.empty(sumL, dtype=torch.int32, device=dev)
        self.nscratch = torch.empty(sumL, dtype=torch.float64, device=dev)
        self.res2 = torch.empty(sumL, dtype=torch.float64, device=dev)
        self.out_dev = torch.empty(2 * sumL + nb + 2, dtype=torch.float64, device=dev)
        self.out_host = torch.empty(2 * sumL + nb + 2, dtype=torch.float64).pin_memory()
        self.maxL, self.maxQ = max(L_), max(Q_)
        # Gram pre-rotation of the projected matrix (gtn_gram_rotate): the Jacobi SVD then runs on hSq
        self.prerotate = PREROTATE and 2 <= self.maxL <= 80
        self.rot_sweeps = torch.zeros(nb, dtype=torch.int32, device=dev)
        hJ = self.hSq if self.prerotate else self.hB
        for b in range(nb):
            parr[b].w_off = ws.off(hJ[b])
        self.pdev = _to_dev_bytes(bytes(parr))
        self.vh_delta = ws.off(self.hVk[0]) - ws.off(hJ[0])
        assert all(ws.off(v) - ws.off(h) == self.vh_delta for v, h in zip(self.hVk, hJ))
        self.persistent = 2 <= self.maxL <= PERSISTENT_MAX_ROWS
        self.graphable = self.persistent and WHITEN == "chol"
        self.graphs, self.graph_launches = {}, {}
        self.warmed = set()             # iteration counts whose schedule has run eagerly (its launch tables are cached)
        self.graph_failures = 0
        self.host_sweeps = None
        self.pending = None
        self.epoch = 0                  # number of load() calls: identifies the run whose state the workspace holds
        torch.cuda.current_stream().synchronize()       # metadata uploads done before any capture

    # ---- launch sequences (enqueue only) -------------------------------------------------------
    def load(self, mats):
        self.epoch += 1
        for b in range(self.nb):
            self.ws.view(self.hW[b]).copy_(mats[b])
        _ws_ctranspose(self.ws, list(zip(self.hW, self.hWh)))

    def _whiten(self, slot, rel_thr=1e-13):
        ws, nb = self.ws, self.nb
        code = dtype_code(self.dt)
        if WHITEN == "eigh":
            with prof_region("small_eigh", 1):
                check(lib.gtn_small_eigh_whiten(_ptr(ws.buf), _ptr(ws.buf), code, _ptr(self.g_off), _ptr(self.t_off),
                                                _ptr(self.n_dev), nb, self.maxL, rel_thr, _ptr(self.kept[slot]),
                                                _ptr(self.evals), _ptr(self.e_off), _stream()), "gtn_small_eigh_whiten")
        else:
            # algorithmic model: pivoted Cholesky n^3/3 + inverse of the factor n^3/3 multiply-adds per matrix (8 real
            # flops each for complex128, 2 for float64); the NS Gram slices read once, T written once
            fl = (8 if self.dt == torch.complex128 else 2) * sum(2 * l ** 3 // 3 for l in self.L_)
            by = ws.buf.element_size() * sum((self.NS + 1) * l * l for l in self.L_)
            with prof_region("chol_whiten", 1, by, fl):
                check(lib.gtn_chol_whiten(_ptr(ws.buf), _ptr(ws.buf), code, _ptr(self.g_off), _ptr(self.t_off),
                                          _ptr(self.n_dev), nb, self.maxL, self.NS, rel_thr, _ptr(self.kept[slot]),
                                          _ptr(self.chol_scratch) if self.chol_scratch is not None else None,
                                          _stream()), "gtn_chol_whiten")

    def _gram_done(self, side):
        """hook between the Gram GEMM and the whitening (column-sharded plans complete the q-side sums here)"""

    def _gram(self, cur, curH):
        """hT1_b = sum_s cur_b[:, Ks] cur_b[:, Ks]^H  as NS partial slices (one grouped launch).
        curH is None: the GEMM reads its B operand as the conjugate transpose of cur itself
        (GTN_GEMM_B_CONJ_TRANS) -- no transposed copy, one launch fewer per orthonormalisation."""
        ws, NS = self.ws, self.NS
        if NS == 1 and curH is not None:
            _ws_gemm(ws, list(zip(cur, curH, self.hT1)))
            return
        groups = []
        for i, (a, c) in enumerate(zip(cur, self.hT1)):
            _, l, K = ws.items[a]
            kc = -(-K // NS)
            for sp in range(NS):
                k0 = min(sp * kc, K)
                kk = min(kc, K - k0)
                if curH is None:
                    groups.append(dict(a_off=ws.off(a) + k0, b_off=ws.off(a) + k0, c_off=ws.off(c) + sp * l * l,
                                       lda=K, ldb=K, ldc=l, m=l, n=l, k=kk, alpha=1.0, beta=0.0, flags=1))
                    continue
                bh = curH[i]
                groups.append(dict(a_off=ws.off(a) + k0, b_off=ws.off(bh) + k0 * l, c_off=ws.off(c) + sp * l * l,
                                   lda=K, ldb=l, ldc=l, m=l, n=l, k=kk, alpha=1.0, beta=0.0))
        key = ("wsgram", str(ws.dtype), tuple(tuple(sorted(g.items())) for g in groups))
        _cached(key, lambda: GemmPlan(groups, ws.dtype)).run(ws.buf, ws.buf, ws.buf)

    def orth(self, src, dst, side, passes, robust=False):
        ws = self.ws
        hC = self.hCp if side == "p" else self.hCq
        hS = self.hSp if side == "p" else self.hSq
        # robust: SHIFTED Cholesky QR (sCholQR3 with two shifted passes).  A panel whose singular values span more than
        # 3e-7 has a Gram matrix that plain pivoted Cholesky truncates (the directions below sqrt(1e-13) s_0 are
        # dropped: the third decomposition of an ATRG step at chi >= 128 has s_k ~ 1e-8 s_0).  Adding 1e-9 trace(G) to
        # the diagonal makes the factorisation exist for any range, T X then has condition <= 3e4 per pass with every
        # direction kept (scaled, not cut), and two plain passes finish the job: 4 passes of 4 launches instead of the
        # one-sided Jacobi SVD of the l x p panel (17 ms at l = 128, and an all-gather when the panel is sharded).
        shifted = 2 if robust else 0
        total = passes + shifted
        tmp = self.hYh if side == "p" else self.hZh                 # second intermediate (the source panel's own storage
        cur = src                                                   # is free once the first pass has consumed it)
        for ps in range(total):
            if GEMM_CONJ_TRANS and not robust:
                self._gram(cur, None)                                # Gram  l x l  straight from cur
            else:
                _ws_ctranspose(ws, list(zip(cur, hC)))
                self._gram(cur, hC)
            self._gram_done(side)
            if ps < shifted:
                check(lib.gtn_gram_shift(_ptr(ws.buf), dtype_code(self.dt), _ptr(self.g_off), _ptr(self.n_dev), self.nb,
                                         self.NS, ROBUST_SHIFT, _stream()), "gtn_gram_shift")
                count()
            self._whiten(0 if ps == 0 else 1, rel_thr=1e-15 if ps < shifted else 1e-13)
            if ps == total - 1:
                out = dst
            else:
                out = hS if cur is not hS else tmp
            _ws_gemm(ws, list(zip(self.hT2, cur, out)))
            cur = out

    def start(self, passes, robust=False):
        _ws_gemm(self.ws, list(zip(self.hG, self.hWh, self.hYh)))
        self.orth(self.hYh, self.hQh, "p", passes, robust)

    def iterate(self, last, robust=False):
        ws = self.ws
        _ws_gemm(ws, list(zip(self.hQh, self.hW, self.hZh)))
        # (skipping this re-orthonormalisation -- one whitening per iteration -- was tried: the iterate's dynamic
        # range squares, directions below ~1e-4 s_0 drop out of the Gram matrix, and every site of the Z2 TRG /
        # ATRG chains eventually stalled)
        self.orth(self.hZh, self.hPh, "q", 1, robust)
        _ws_gemm(ws, list(zip(self.hPh, self.hWh, self.hYh)))
        self.orth(self.hYh, self.hQh, "p", 2 if last else 1, robust)

    def check_enqueue(self, allow_host=True):
        """B = Qh W, its Jacobi SVD, the Ritz vectors and the residual norms; results into out_host."""
        ws, nb, dev, dt = self.ws, self.nb, self.dev, self.dt
        code = dtype_code(dt)
        st = _stream()
        _ws_gemm(ws, list(zip(self.hQh, self.hW, self.hB)))
        Wp = _ptr(ws.buf)
        if self.prerotate:
            # B' = T B, Qh' = T Qh with the unitary T from the Gram matrix of B: rows of B' nearly orthogonal
            if GEMM_CONJ_TRANS:
                self._gram(self.hB, None)
            else:
                _ws_ctranspose(ws, list(zip(self.hB, self.hCq)))
                self._gram(self.hB, self.hCq)
            # (Cholesky part only: n^3/3 multiply-adds; the shared-memory Jacobi sweeps on the factor are data dependent)
            fl = (8 if dt == torch.complex128 else 2) * sum(l ** 3 // 3 for l in self.L_)
            by = ws.buf.element_size() * sum((self.NS + 1) * l * l for l in self.L_)
            with prof_region("gram_rotate", 1, by, fl):
                check(lib.gtn_gram_rotate(Wp, Wp, code, _ptr(self.g_off), _ptr(self.t_off), _ptr(self.n_dev), nb,
                                          self.maxL, self.NS, 1e-15, ROTATE_TOL, 30, _ptr(self.rot_sweeps), st),
                      "gtn_gram_rotate")
            _ws_gemm(ws, list(zip(self.hT2, self.hB, self.hSq)))
            _ws_gemm(ws, list(zip(self.hT2, self.hQh, self.hSp)))
            a0, a1 = ws.off(self.hQh[0]), ws.off(self.hSp[0])
            nq = sum(l * p for l, p in zip(self.L_, self.P_))
            ws.buf[a0: a0 + nq].copy_(ws.buf[a1: a1 + nq])
        check(lib.gtn_jacobi_init(Wp, Wp, code, _ptr(self.pdev), nb, self.maxL, _ptr(self.rn2), _ptr(self.fro2),
                                  _ptr(self.rn_off), st), "gtn_jacobi_init")
        count()
        self.host_sweeps = None
        done = False
        P = (self.maxL + 1) & ~1
        esz = ws.buf.element_size()
        rb = sum(2 * esz * (l * q + l * l) for l, q in zip(self.L_, self.Q_))
        if self.persistent:
            self._jbytes = rb * (P - 1)              # per sweep; scaled by the sweep count in read()
            with prof_region("jacobi_persistent", 1, 0):
                rc = lib.gtn_jacobi_persistent(Wp, Wp, code, _ptr(self.pdev), nb, self.maxL, JACOBI_TOL, _ptr(self.offd),
                                               _ptr(self.rn2), _ptr(self.fro2), _ptr(self.rn_off), JACOBI_MAX_SWEEPS,
                                               _ptr(self.sw), st, JACOBI_EARLY_STOP)
            if rc == 0:
                done = True
            elif rc == -2:
                self.persistent = self.graphable = False
            else:
                check(rc, "gtn_jacobi_persistent")
        if not done:
            if not allow_host:
                raise _cabi.GtnError("persistent Jacobi kernel unavailable inside a captured schedule")
            sweeps = 0
            self.offd.zero_()
            while True:
                with prof_region("jacobi_round", P - 1, rb * (P - 1)):
                    check(lib.gtn_jacobi_sweep(Wp, Wp, code, _ptr(self.pdev), nb, self.maxL, self.maxQ, JACOBI_TOL,
                                               _ptr(self.offd), _ptr(self.rn2), _ptr(self.fro2), _ptr(self.rn_off), st),
                          "gtn_jacobi_sweep")
                sweeps += 1
                if float(self.offd[:nb].max().item()) <= JACOBI_TOL ** 2:
                    break
                if sweeps >= JACOBI_MAX_SWEEPS:
                    raise _cabi.GtnError("Jacobi SVD did not converge in %d sweeps" % sweeps)
            self.host_sweeps = sweeps
        vh_ptr = C.c_void_p(ws.buf.data_ptr() + self.vh_delta * esz)
        check(lib.gtn_jacobi_finish(Wp, Wp, Wp, vh_ptr, _ptr(self.s_dev), code, _ptr(self.pdev), _ptr(self.odev),
                                    _ptr(self.order), _ptr(self.nscratch), nb, self.maxL, self.maxQ, st),
              "gtn_jacobi_finish")
        count(2)
        _ws_ctranspose(ws, list(zip(self.hUb, self.hUbH)))
        _ws_gemm(ws, list(zip(self.hUbH, self.hQh, self.hUh)))          # Uh = Ub^H Qh  (l x p)
        # certificate rows:  Eh_i = v_i^H W^H - s_i u_i^H   (l x p)
        _ws_gemm(ws, list(zip(self.hVk, self.hWh, self.hXh)))
        for b in range(nb):
            ws.view(self.hD[b]).diagonal().copy_(self.s_dev[self.soff[b]: self.soff[b] + self.L_[b]])
        _ws_gemm(ws, list(zip(self.hD, self.hUh, self.hXh)), alpha=-1.0, beta=1.0)
        for b in range(nb):
            x = ws.view(self.hXh[b])
            with prof_region("row_sumsq", 1, x.numel() * x.element_size()):
                check(lib.gtn_row_sumsq(_ptr(x), _ptr(self.res2[self.soff[b]:]), self.L_[b], self.P_[b], code, st),
                      "gtn_row_sumsq")
        sL = self.sumL
        self.out_dev[:sL].copy_(self.s_dev)
        self.out_dev[sL: 2 * sL].copy_(self.res2)
        self.out_dev[2 * sL: 2 * sL + nb].copy_(self.kept[0])
        self.out_dev[2 * sL + nb:].copy_(self.sw[:2])
        self.out_host.copy_(self.out_dev, non_blocking=True)

    def read(self):
        torch.cuda.current_stream().synchronize()
        o = self.out_host.numpy()
        sL, nb = self.sumL, self.nb
        svals = [o[a: a + l].copy() for a, l in zip(self.soff, self.L_)]
        res = np.sqrt(np.maximum(o[sL: 2 * sL], 0.0))
        kept = o[2 * sL: 2 * sL + nb].astype(np.int64)
        if self.host_sweeps is None:
            sweeps, conv = int(o[2 * sL + nb]), int(o[2 * sL + nb + 1])
            if not conv:
                raise _cabi.GtnError("Jacobi SVD did not converge in %d sweeps" % sweeps)
            if PROF.enabled:
                for i in range(len(PROF.records) - 1, -1, -1):
                    r = PROF.records[i]
                    if r[0] == "jacobi_persistent":
                        if r[4] == 0:
                            PROF.records[i] = r[:4] + (self._jbytes * max(sweeps, 1),) + r[5:]
                        break
        else:
            sweeps = self.host_sweeps
        batched_svd.last_sweeps = sweeps
        return svals, res, kept

    def finalize(self):
        ws = self.ws
        _ws_ctranspose(ws, list(zip(self.hUh, self.hU)))
        return [(ws.view(self.hU[b]), None, ws.view(self.hVk[b])) for b in range(self.nb)]

    _allreduce = None                   # (column-sharded plans complete the sums over columns across ranks)

    def rank_certificate(self, b, nD, thr):
        """bound of || W_b - U_D U_D^H W_b ||_2 from the Ritz vectors of the last check (deflated_norm_bound)"""
        ws = self.ws
        _ws_ctranspose(ws, [(self.hUh[b], self.hU[b])])
        RANK_CHECK_STATS["calls"] += 1
        bound = deflated_norm_bound(ws.view(self.hW[b]), ws.view(self.hUh[b]), ws.view(self.hU[b]), self.L_[b], nD, thr,
                                    self._allreduce)
        RANK_CHECK_STATS["certified"] += int(bound <= thr)
        return bound

    # ---- CUDA graph of the steady-state schedule -----------------------------------------------
    def schedule(self, n):
        self.start(2 if n == 0 else 1)
        for it in range(1, n + 1):
            self.iterate(it == n)
        self.check_enqueue(allow_host=False)

    def gkey(self, n):
        return n

    def graph(self, n):
        n_, n = n, self.gkey(n)
        g = self.graphs.get(n)
        if g is None:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            c0 = _cabi.launch_count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                g.capture_begin()
                try:
                    self.schedule(n_)
                finally:
                    g.capture_end()
            cur.wait_stream(side)
            self.graph_launches[n] = _cabi.launch_count - c0
            _cabi.launch_count = c0                      # captured, not executed
            self.graphs[n] = g
        return g


def _trunc_plan(key, P_, Q_, ks, L_, dt, dev):
    plan = _trunc_plans.get(key)
    if plan is None:
        plan = _TruncPlan(P_, Q_, ks, L_, dt, dev)
        plan.cached = plan.nbytes <= TRUNC_PLAN_CACHE_BYTES
        if plan.cached:
            if len(_trunc_plans) >= 32:      # (speculative callers hold one plan per call site: ATRG alone has six)
                _trunc_plans.pop(next(iter(_trunc_plans)))
            _trunc_plans[key] = plan
    return plan


def _trunc_certificate(svals, res, kept_host, ks, L_, rank_check=None):
    """(ok, worst, reject) for the Ritz values `svals`, residual norms `res` and whitening ranks `kept_host`
    of one check: ok = every kept triplet has residual <= TRUNC_TOL * s_0 AND the kept ones are the largest;
    reject = the subspace lost directions while fewer than k triplets were found (count not trustworthy).
    rank_check(b, nD, thr) -> bound of the deflated matrix's norm (plan.rank_certificate): decides the count of
    non-zero values of a numerically rank-deficient sector; svals[b] is then lowered in place beyond nD."""
    ok, o, worst = True, 0, 0.0
    for b in range(len(svals)):
        s = svals[b]
        s0 = s[0] if len(s) else 0.0
        nnz = int(np.sum(np.abs(s / (abs(s0) + 1e-14)) > 1e-14)) if len(s) else 0
        kk = min(ks[b], nnz)
        # (a) the smallest value this sector would keep lies below the resolution of the subspace iteration (its
        #     null-space noise is ~1e-14 s_0, LAPACK's ~1e-16): whether it passes the reference's rank rule
        #     s_i / s_0 > 1e-14 cannot be read off the Ritz values (seen on the Z2 chain at chi = 64: exact rank 16
        #     per sector, 19 "non-zero" Ritz values);
        # (b) fewer triplets than requested AND the Gram whitening could not resolve every direction: a
        #     small-but-valid singular direction may have been dropped.
        # Both are settled by the norm of the deflated matrix; without that certificate the full SVD decides.
        band = kk > 0 and s[kk - 1] <= RANK_NOISE * s0
        dropped = nnz < ks[b] and kept_host is not None and kept_host[b] < L_[b] and nnz >= kept_host[b]
        if band or dropped:
            nD = int(np.sum(s > RANK_NOISE * s0))
            thr = 1e-14 * (abs(s0) + 1e-14)
            if rank_check is None or nD == 0 or nD >= len(s):
                if DEBUG_TRUNC:
                    print("[trunc] sector", b, "rank undecidable here: band", band, "dropped", dropped, "nD", nD, flush=True)
                return False, worst, True
            bound = rank_check(b, nD, thr)
            if DEBUG_TRUNC:
                print("[trunc] sector", b, "rank certificate: band", band, "dropped", dropped, "nD", nD, "of k", ks[b],
                      "bound/s0 %.2e" % (bound / max(s0, 1e-300)), "s[nD-1:nD+3]/s0",
                      np.array2string(s[max(nD - 1, 0): nD + 3] / max(s0, 1e-300), precision=2), flush=True)
            if not bound <= thr:
                return False, worst, True
            s[nD:] = np.minimum(s[nD:], bound)
            nnz = nD
            kk = min(ks[b], nnz)
        if kk > 0:
            worst = max(worst, float(np.max(res[o: o + kk])) / max(s0, 1e-300))
        if kk > 0 and np.max(res[o: o + kk]) > TRUNC_TOL * s0:
            ok = False
        if 0 < kk < len(s):
            # the accepted triplets must also be the LARGEST ones: a Ritz pair j > kk with residual
            # r_j stands for an exact singular value within r_j of it, so none of them may reach
            # above the smallest accepted value (a direction that is still poorly represented in
            # the subspace shows up here as a low Ritz value with a large residual)
            hi = s[kk:] + res[o + kk: o + len(s)]
            if np.max(hi) > s[kk - 1] + max(TRUNC_TOL * s0, res[o + kk - 1]):
                ok = False
                worst = max(worst, float(np.max(hi) - s[kk - 1]) / max(s0, 1e-300))
        o += L_[b]
    return ok, worst, False


ADAPT = [True]              # False: iteration counts are frozen (gauge2d.freeze): no lowering, no probing
PROBE_EVERY = int(__import__("os").environ.get("GTN_PROBE_EVERY", "6"))
_trunc_probe = {}           # (batch shape, call site) -> {"stable": accepts without change, "every": probe interval, ...}


def _trunc_accept(key, it, worst, spare=1, clean=True):
    """bookkeeping of an accepted run: remember the iteration count for this (shape, call site).
    * lowered when the run passed with a margin of d convergence factors (d - spare fewer next time, at most half;
      speculative runs keep two spare factors: a failed speculation costs the tail of the caller's step);
    * the residuals floor at rounding level, so a margin cannot show that far fewer iterations would do (a chain
      whose spectrum gets easier kept its early count of 5 where 1 suffices): after `every` accepts of the same
      count the memory is dropped and the next run derives the count afresh (check after the range finder, then
      at the predicted iteration).  When that does not find a lower count, probing becomes four times rarer;
    * a clean accept (passed at its first check) never raises the count: a whole-step graph keeps replaying its
      recorded count while a lower one waits for the graph to be recorded again."""
    if not ADAPT[0] and clean and key in _trunc_iters_hint:
        _trunc_fail[key] = 0
        return
    rate = _trunc_rate.get(key, 0.2)
    d = int(math.log(max(worst, 1e-16) / TRUNC_TOL) / math.log(rate)) if worst < TRUNC_TOL else 0
    new = max(it - min(max(d - spare, 0), (it + 1) // 2), 0)
    st = _trunc_probe.setdefault(key, {"stable": 0, "every": PROBE_EVERY, "before": None})
    old = _trunc_iters_hint.get(key)
    _trunc_fail[key] = 0
    if old is None and st["before"] is not None:
        # the run after a probe: did deriving the count afresh pay?
        if new >= st["before"]:
            st["every"] = min(st["every"] * 4, 1 << 20)
        st["before"], st["stable"] = None, 0
    elif clean:
        if old is not None:
            new = min(new, old)
        st["stable"] = st["stable"] + 1 if (new == old and new == it) else 0      # this very count keeps passing
        if PROBE_EVERY > 0 and st["stable"] >= st["every"] and new > 0:
            st["before"], st["stable"] = new, 0
            _trunc_iters_hint.pop(key, None)
            return
    else:
        st["stable"] = 0
    _trunc_iters_hint[key] = new


class SpeculativeSVD:
    """Result of a truncated SVD whose certificate has NOT been read back yet (the steady-state schedule
    was replayed as a CUDA graph and nothing was synchronised).  `outs[b] = (U, s_dev, Vh)` are device
    views valid until the next run of the same batch shape; the caller may enqueue dependent work and must
    call verify() before trusting any of it: verify() synchronises, evaluates the certificate and the
    assumption that every sector has at least k_b non-zero singular values, and returns True/False."""

    def __init__(self, plan, key, it, ks, L_, pkey=None):
        self.plan, self.key, self.it, self.ks, self.L_, self.pkey = plan, key, it, ks, L_, pkey
        fin = plan.finalize()
        self.outs = [(u, plan.s_dev[o: o + l], v) for (u, _, v), o, l in zip(fin, plan.soff, plan.L_)]
        self.svals = None
        self.readback = None
        self.ok = None
        self.epoch = plan.epoch
        plan.pending = self

    def discard(self):
        """Drop a run whose INPUT turned out to be unverified (an earlier decomposition of the same step failed
        its certificate): no read-back, no bookkeeping."""
        if self.ok is None:
            self.ok = False
            if self.plan.pending is self:
                self.plan.pending = None

    def reverify(self):
        """certificate of a NEW execution of the same enqueued schedule (whole-step graph replay)"""
        self.ok = None
        self.plan.pending = self
        return self.verify()

    def verify(self):
        if self.ok is not None:
            return self.ok
        plan = self.plan
        plan.pending = None
        try:
            svals, res, kept_host = plan.read()
        except _cabi.GtnError:
            self.ok = False
            return False
        ok, worst, reject = _trunc_certificate(svals, res, kept_host, self.ks, self.L_)
        full = all(int(np.sum(np.abs(s / (abs(s[0]) + 1e-14)) > 1e-14)) >= k for s, k in zip(svals, self.ks))
        truncated_svd_batch.last_iters = self.it
        if DEBUG_TRUNC:
            print("[trunc] speculative it", self.it, "worst %.2e" % worst, "ok", ok, "full rank", full, flush=True)
        self.svals = svals
        self.readback = (svals, res, kept_host)
        self.ok = bool(ok and full and not reject)
        if self.ok and FORCE_VERIFY_FAIL[0] is not None and FORCE_VERIFY_FAIL[0](self):
            self.ok = False             # test hook: exercise the callers' mis-speculation paths on a sound result
            return False
        if self.ok:
            _trunc_accept(self.key, self.it, worst, spare=2)
        return self.ok


ONE_CALL = bool(int(__import__("os").environ.get("GTN_ONE_CALL", "1")))
ONE_CALL_STATS = {"calls": 0, "accepted": 0}


def _onecall_arrays(P_, Q_, ks):
    nb = len(P_)
    return (C.c_int64 * nb)(*P_), (C.c_int64 * nb)(*Q_), (C.c_int32 * nb)(*ks)


def _onecall_bytes(P_, Q_, ks, dt):
    m_, n_, k_ = _onecall_arrays(P_, Q_, ks)
    return int(lib.gtn_workspace_bytes(_cabi.GTN_OP_SECTOR_SVD_TRUNC, dtype_code(dt), len(P_), m_, n_, k_))


def _truncated_onecall(mats, ks, key, P_, Q_, robust=False):
    """truncated_svd_batch through the one-call C ABI (include/gtn_b200.h: gtn_sector_svd_trunc): subspace iteration,
    certificate checks and rank certificate run inside the library; returns the same [(U, s, Vh)] or None."""
    dev, dt = mats[0].device, mats[0].dtype
    nb, code = len(mats), dtype_code(dt)
    m_, n_, k_ = _onecall_arrays(P_, Q_, ks)
    nbytes = int(lib.gtn_workspace_bytes(_cabi.GTN_OP_SECTOR_SVD_TRUNC, code, nb, m_, n_, k_))
    if nbytes <= 0:
        raise _cabi.GtnError("gtn_workspace_bytes failed with status %d" % nbytes)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    src = [m.contiguous() for m in mats]
    U = [torch.empty(p, k, dtype=dt, device=dev) for p, k in zip(P_, ks)]
    Vh = [torch.empty(k, q, dtype=dt, device=dev) for q, k in zip(Q_, ks)]
    ptrs = lambda ts: (C.c_void_p * nb)(*[t.data_ptr() for t in ts])
    S = (C.c_double * sum(ks))()
    rank = (C.c_int32 * nb)()
    info = _cabi.SvdInfo()
    info.robust = 1 if robust else 0
    info.start_iters = int(_trunc_iters_hint.get(key) or 0)
    info.rate = float(_trunc_rate.get(key, 0.0))
    ONE_CALL_STATS["calls"] += 1
    rc = lib.gtn_sector_svd_trunc(ptrs(src), m_, n_, nb, code, k_, 1e-14, ptrs(U), S, ptrs(Vh), rank, _ptr(ws), nbytes,
                                  C.byref(info), _stream())
    count(int(info.launches))
    truncated_svd_batch.last_iters = int(info.iters)
    batched_svd.last_sweeps = int(info.sweeps)
    if DEBUG_TRUNC:
        print("[trunc one-call] rc", rc, "iters", info.iters, "checks", info.checks, "worst %.2e" % info.worst,
              "start", info.start_iters, "launches", info.launches, flush=True)
    if rc == _cabi.GTN_ERR_NOT_CONVERGED:
        return None
    check(rc, "gtn_sector_svd_trunc")
    ONE_CALL_STATS["accepted"] += 1
    if info.rate > 0:
        _trunc_rate[key] = float(info.rate)
    _trunc_accept(key, int(info.iters), float(info.worst), spare=1, clean=(info.checks == 1))
    out, o = [], 0
    sv = np.frombuffer(S, dtype=np.float64).copy()
    for b in range(nb):
        out.append((U[b], sv[o: o + ks[b]], Vh[b]))
        o += ks[b]
    return out


def truncated_svd_batch(mats, ks, robust=False, speculative=False, resume=None):
    """Top-k_b singular triplets of every matrix in `mats` by randomized subspace iteration:
         Yh = G Wh ; Qh = orth_rows(Yh) ; [Zh = Qh W ; Ph = orth_rows(Zh) ; Yh = Ph Wh ; Qh = orth_rows(Yh)]*
         B = Qh W (l x q) ; B = Ub S Vh (small one-sided Jacobi) ; U = Qh^H Ub
    All products run on the DMMA GEMM kernel.  orth_rows whitens with the l x l Gram matrix
    (gtn_chol_whiten, one CTA per matrix; two passes for the final basis) -- or, with
    robust=True, runs the one-sided Jacobi kernels on the l rows (keeps directions that a Gram matrix
    cannot resolve).  The final small SVD of B always uses one-sided Jacobi, so the kept singular
    values do not go through a Gram matrix.  The result is accepted only with a certificate: the
    residuals || W v_i - s_i u_i || of the first min(k, rank) triplets are <= TRUNC_TOL * s_0.
    Returns None when the certificate fails after TRUNC_MAX_ITERS refinements, or when the Gram
    whitening dropped directions while fewer than k triplets were found -- the caller then retries
    with robust=True or runs the full Jacobi SVD.
    The returned U / Vh are views into the plan's workspace: valid until the next call with the same
    batch shape (callers unpack them into tensors straight away).
    Replaces LAPACK's full gesdd in reference SortedSVD (__init__.py:3932) when only the first
    `cutoff` singular triplets are kept (:3943-3948)."""
    dev, dt = mats[0].device, mats[0].dtype
    nb = len(mats)
    P_ = [m.shape[0] for m in mats]
    Q_ = [m.shape[1] for m in mats]
    L_ = [min(p, q, subspace_rows(k), TRUNC_LMAX) for p, q, k in zip(P_, Q_, ks)]
    pkey = (tuple(P_), tuple(Q_), tuple(ks), str(dt), str(dev))
    key = (pkey, SVD_SITE[0])               # iteration hints / failure memory are per call site
    if robust:
        key = key + ("robust",)
    elif _trunc_robust.get(key) and resume is None and not CAPTURING_STEP[0]:
        # this site's spectrum spans more than the Gram whitening resolves: straight to the shifted-Cholesky-QR run
        # (also for speculative callers: they accept a verified, non-speculative result; measured on the chi = 128
        # ATRG chain: the plain attempt + its rank certificate cost 135 ms of every 620 ms step before being rejected)
        return truncated_svd_batch(mats, ks, robust=True)
    if (ONE_CALL and resume is None and not CAPTURING_STEP[0] and not PROF.enabled
            and _onecall_bytes(P_, Q_, ks, dt) > TRUNC_PLAN_CACHE_BYTES):
        # large sectors (chi >= 128: no cached workspace, no recorded graph): the whole driver loop runs behind
        # ONE C-ABI call (gtn_sector_svd_trunc, csrc/gtn_sector.cu); this host only keeps the iteration memory
        return _truncated_onecall(mats, ks, key, P_, Q_, robust)
    fails = _trunc_fail.get(key, 0)
    if fails >= 2 and not robust:
        # this shape keeps failing the certificate (flat spectrum): go straight to the full SVD, but
        # re-try every 16th call in case the spectrum has changed
        _trunc_fail[key] = fails + 1 if fails < 17 else 1
        return None
    # speculative callers get a workspace of their own per call site: the three same-shape SVDs of an ATRG step
    # are then in flight behind each other without waiting for one another's certificate
    if resume is not None and resume.pkey == pkey and not robust:
        plan = resume.plan
    else:
        plan = _trunc_plan((pkey, SVD_SITE[0]) if speculative else pkey, P_, Q_, ks, L_, dt, dev)
    if getattr(plan, "pending", None) is not None:
        plan.pending.verify()               # an unverified speculative run still owns the read-back buffer
    resumed = (resume is not None and resume.plan is plan and resume.readback is not None
               and resume.epoch == plan.epoch and not robust)
    if resumed:
        # continue a speculative run whose certificate failed: the workspace still holds its iterate, the
        # check has been read back already; bookkeeping goes to the speculative call site
        key = resume.key
    else:
        plan.load(mats)
    hint = _trunc_iters_hint.get(key)
    # steady state: as many iterations as the last accepted run needed, one fewer when that run passed
    # with a margin of one iteration's convergence factor (a failed check costs ~4 iterations' worth).
    start_it = hint or 0
    replayed = False
    if CAPTURING_STEP[0]:
        # the caller records its whole step as ONE CUDA graph: the schedule 'start, n iterations, check' is enqueued
        # inline (no nested replay, no read-back); the caller verifies the certificate after every replay
        if not (speculative and hint is not None and not robust and not resumed and plan.cached and plan.graphable
                and (plan.gkey(start_it) in plan.graphs or plan.gkey(start_it) in plan.warmed)):
            raise NotCapturable("truncated SVD of shape %s is not in its steady state" % (pkey[:3],))
        plan.schedule(start_it)
        return SpeculativeSVD(plan, key, start_it, ks, L_, pkey)
    if resumed:
        replayed, start_it = True, resume.it
    elif (USE_GRAPHS and hint is not None and not robust and plan.cached and plan.graphable and not PROF.enabled
          and (plan.gkey(start_it) in plan.graphs or plan.gkey(start_it) in plan.warmed)):
        # a schedule is recorded only after it has run eagerly once: building a launch table uploads it from
        # pageable host memory, which a capturing stream refuses
        try:
            g = plan.graph(start_it)
            g.replay()
            count(plan.graph_launches[plan.gkey(start_it)])
            replayed = True
        except (RuntimeError, _cabi.GtnError) as exc:
            if DEBUG_TRUNC:
                print("[trunc] CUDA graph capture/replay failed, falling back to eager launches:", repr(exc)[:300], flush=True)
            plan.graph_failures += 1
            plan.warmed.discard(plan.gkey(start_it))
            if plan.graph_failures >= 3:
                plan.graphable = False
            plan.graphs.pop(plan.gkey(start_it), None)
            torch.cuda.synchronize()
    if speculative and replayed:
        return SpeculativeSVD(plan, key, start_it, ks, L_, pkey)
    prev_worst, prev_it, next_check = None, None, 0
    for it in range(TRUNC_MAX_ITERS + 1):
        if replayed and it <= start_it:
            if it < start_it:
                continue                     # done inside the graph, check included
        else:
            if it == 0:
                plan.start(2 if start_it == 0 else 1, robust)
            else:
                plan.iterate(it >= start_it and it >= next_check, robust)
            if it < start_it or it < next_check:
                continue
            plan.check_enqueue()
            if it == start_it and not robust:
                plan.warmed.add(plan.gkey(start_it))        # exactly the launches of schedule(start_it) have now been built
        if resumed and it == start_it:
            svals, res, kept_host = resume.readback
        else:
            try:
                svals, res, kept_host = plan.read()
            except _cabi.GtnError:
                # the small Jacobi SVD of the projected matrix did not converge: let the caller run the full SVD
                _trunc_fail[key] = _trunc_fail.get(key, 0) + 1
                return None
        if robust:
            kept_host = None
        ok, worst, reject = _trunc_certificate(svals, res, kept_host, ks, L_, rank_check=plan.rank_certificate)
        if reject:
            truncated_svd_batch.last_iters = it
            return None
        truncated_svd_batch.last_iters = it
        if DEBUG_TRUNC:
            print("[trunc] it", it, "worst %.2e" % worst, "ok", ok, "graph", replayed, "shapes", list(zip(P_, Q_)), "k", ks,
                  "L", L_, "s[k-2:k+3]/s0",
                  [np.array2string(sv[max(0, k - 2):k + 3] / sv[0], precision=5) for sv, k in zip(svals, ks)], flush=True)
        if prev_worst is not None and prev_worst > 0 and worst > 0 and it > prev_it:
            _trunc_rate[key] = min(max((worst / prev_worst) ** (1.0 / (it - prev_it)), 1e-3), 0.9)
        if not ok and prev_worst is not None and it >= 2:
            rate = (worst / prev_worst) ** (1.0 / max(it - prev_it, 1)) if prev_worst > 0 else 1.0
            if rate > 0.6:
                # the subspace iteration has stalled (flat spectrum at the cut): stop wasting GEMMs
                _trunc_fail[key] = _trunc_fail.get(key, 0) + 1
                return None
            # geometric convergence: skip the (Jacobi-SVD + residual) check until the predicted
            # iteration, it costs more than the two GEMM + two whitening launches of an iteration
            need = math.log(max(TRUNC_TOL * 0.3, 1e-300) / max(worst, 1e-300)) / math.log(max(rate, 1e-3))
            next_check = it + max(1, min(int(math.ceil(need)), 6))
            if next_check > TRUNC_MAX_ITERS:
                _trunc_fail[key] = _trunc_fail.get(key, 0) + 1
                return None
        if not ok and prev_worst is None and key in _trunc_rate and worst > 0:
            # first failed check of this call: schedule the next one with the convergence factor measured on
            # earlier calls from this site instead of checking again after a single iteration
            need = math.log(max(TRUNC_TOL * 0.3, 1e-300) / worst) / math.log(_trunc_rate[key])
            next_check = min(it + max(1, min(int(math.ceil(need)), 6)), TRUNC_MAX_ITERS)
        prev_worst, prev_it = worst, it
        if ok:
            _trunc_accept(key, it + (1 if resumed else 0), worst, spare=2 if resumed else 1,
                          clean=(it == start_it and not resumed))
            out = plan.finalize()
            return [(u, svals[b], v) for b, (u, _, v) in enumerate(out)]
    _trunc_fail[key] = _trunc_fail.get(key, 0) + 1
    return None


truncated_svd_batch.last_iters = 0
