"""TEST DOUBLE for the CUDA library (tests only; nothing under grassmanntn_b200/ imports this).

`install(monkeypatch)` lets the product's HOST logic -- planner, launch tables, GEMM group lists, decomposition
packing, rank rule, unpacking -- run without a GPU by standing in for the C ABI with numpy:
  gtn_sign_permute      -> the numpy emulation of the kernel's addressing / sign rule (test_tables_cpu.emulate)
  gtn_grouped_gemm      -> numpy matmul per group (offsets, leading dimensions, batch strides, alpha/beta, conj-trans B)
  gtn_jacobi_* (full)   -> numpy.linalg.svd behind _engine.batched_svd's interface
  gtn_sumsq / rowsum / row_sumsq / dot / pow_rcond / scale -> numpy on the host buffers
  install(truncated=True) additionally: gtn_chol_whiten (pivoted Cholesky), gtn_gram_rotate (eigenvectors of the
  Gram matrix), gtn_jacobi_init / _persistent / _finish (numpy SVD behind the kernels' contracts) -- the randomized
  subspace iteration with its residual certificate then runs its eager schedule on the host
CUDA graphs, speculation and the multi-GPU mode are NOT emulated (GPU tests cover them).
It is not a fallback: the product raises without a CUDA device (test_no_cpu_fallback) unless a test installs this."""
import ctypes as C

import numpy as np
import torch

from test_tables_cpu import emulate


def _arr(ptr, n, dtype):
    addr = ptr.value if isinstance(ptr, C.c_void_p) else int(ptr)
    if n <= 0:
        return np.zeros(0, dtype=dtype)
    nd = n * (2 if dtype == np.complex128 else 1)
    raw = np.ctypeslib.as_array((C.c_double * nd).from_address(addr))
    return raw.view(np.complex128) if dtype == np.complex128 else raw


def _popcount(x):
    x = x.astype(np.uint64)
    c = np.zeros(x.shape, dtype=np.uint64)
    for _ in range(32):
        c += x & np.uint64(1)
        x = x >> np.uint64(1)
    return c


def emulate_vec(fields, tabs, src, dst, scale=1.0):
    """the same addressing / sign rule as test_tables_cpu.emulate (csrc/gtn_permute.cu), vectorised over the
    product of the super-axis tables so that chi = 32 tensors are affordable; checked against the scalar
    emulation in tests/test_host_double.py::test_vectorised_emulation_equals_scalar"""
    i = np.full((1,), fields["in_base"], dtype=np.int64)
    o = np.full((1,), fields["out_base"], dtype=np.int64)
    e = np.full((1,), fields["const"], dtype=np.uint64)
    macc = np.zeros((1,), dtype=np.uint64)
    for t in tabs:
        P = t["P"].astype(np.uint64)
        low = P & np.uint64(0x0FFFFFFF)
        i = (i[:, None] + t["in_off"].astype(np.int64)[None, :]).ravel()
        o = (o[:, None] + t["out_off"].astype(np.int64)[None, :]).ravel()
        cross = _popcount(low[None, :] & macc[:, None]) & np.uint64(1)
        e = ((e[:, None] ^ (P >> np.uint64(31))[None, :]) ^ cross).ravel()
        macc = (macc[:, None] ^ t["M"].astype(np.uint64)[None, :]).ravel()
    v = src[i]
    if fields["conj"]:
        v = np.conj(v)
    dst[o] = scale * np.where((e & np.uint64(1)).astype(bool), -v, v)


class HostPlan:
    def __init__(self, jobs):
        self.jobs = jobs

    def run(self, src, dst, scale=1.0):
        s, d = src.numpy(), dst.numpy()
        for f, tabs in self.jobs:
            emulate_vec(f, tabs, s, d, scale)


class HostGemm:
    def __init__(self, groups, dtype, config=None):
        self.groups, self.n, self.config = [dict(g) for g in groups], len(groups), config
        cplx = dtype == torch.complex128
        self.flops = sum((8 if cplx else 2) * g.get("batch", 1) * g["m"] * g["n"] * g["k"] for g in groups)
        self.bytes, self.tiles = 0, len(groups)

    def run(self, A, B, Cm):
        a, b, c = A.numpy().reshape(-1), B.numpy().reshape(-1), Cm.numpy().reshape(-1)
        st = np.lib.stride_tricks.as_strided
        it = a.itemsize
        for g in self.groups:
            m, n, k = g["m"], g["n"], g["k"]
            if m == 0 or n == 0:
                continue
            for bi in range(g.get("batch", 1)):
                ao = g["a_off"] + bi * g.get("bsa", 0)
                bo = g["b_off"] + bi * g.get("bsb", 0)
                co = g["c_off"] + bi * g.get("bsc", 0)
                Am = st(a[ao:], (m, k), (g["lda"] * it, it)) if k else np.zeros((m, 0), a.dtype)
                if g.get("flags", 0) & 1:
                    Bm = st(b[bo:], (n, k), (g["ldb"] * it, it)).conj().T if k else np.zeros((0, n), a.dtype)
                else:
                    Bm = st(b[bo:], (k, n), (g["ldb"] * it, it)) if k else np.zeros((0, n), a.dtype)
                Cv = st(c[co:], (m, n), (g["ldc"] * it, it))
                res = g.get("alpha", 1.0) * (Am @ Bm)
                if g.get("beta", 0.0) != 0.0:
                    res = res + g["beta"] * Cv
                Cv[...] = res


def host_batched_svd(mats):
    out = []
    for M in mats:
        u, s, vh = np.linalg.svd(M.numpy(), full_matrices=False)
        out.append((torch.from_numpy(np.ascontiguousarray(u)), s, torch.from_numpy(np.ascontiguousarray(vh))))
    host_batched_svd.last_sweeps = 0
    return out


host_batched_svd.last_sweeps = 0


class HostLib:
    """the C entry points the host code calls directly, on host buffers; everything else -> the real library"""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    @staticmethod
    def _dt(code):
        from grassmanntn_b200 import _engine as E
        return np.complex128 if code == E.dtype_code(torch.complex128) else np.float64

    def gtn_sumsq(self, x, n, code, out, zero_first, stream):
        o = _arr(out, 1, np.float64)
        if zero_first:
            o[0] = 0.0
        v = _arr(x, n, self._dt(code))
        o[0] += float(np.sum(v.real ** 2 + (v.imag ** 2 if np.iscomplexobj(v) else 0.0)))
        return 0

    def gtn_sumabs(self, x, n, code, out, zero_first, stream):
        o = _arr(out, 1, np.float64)
        if zero_first:
            o[0] = 0.0
        o[0] += float(np.sum(np.abs(_arr(x, n, self._dt(code)))))
        return 0

    def gtn_rowsum(self, x, y, rows, cols, code, stream):
        dt = self._dt(code)
        _arr(y, rows, dt)[:] = _arr(x, rows * cols, dt).reshape(rows, cols).sum(axis=1)
        return 0

    def gtn_sum_slices(self, x, y, n, nslices, code, stream):
        dt = np.complex128 if code == 1 else np.float64
        _arr(y, n, dt)[:] = _arr(x, n * nslices, dt).reshape(nslices, n).sum(axis=0)
        return 0

    def gtn_row_sumsq(self, x, y, rows, cols, code, stream):
        v = _arr(x, rows * cols, self._dt(code)).reshape(rows, cols)
        _arr(y, rows, np.float64)[:] = np.sum(np.abs(v) ** 2, axis=1)
        return 0

    def gtn_dot(self, a, b, n, code, out, partial, nparts, stream):
        dt = self._dt(code)
        _arr(out, 1, dt)[0] = np.sum(_arr(a, n, dt) * _arr(b, n, dt))
        return 0

    def gtn_pow_rcond(self, x, n, code, p, rcond, stream):
        v = _arr(x, n, self._dt(code))
        keep = np.abs(v) > rcond
        with np.errstate(all="ignore"):
            r = np.power(np.where(keep, v, 1.0), p)
        v[:] = np.where(keep, r, 0.0)
        return 0

    def gtn_scale(self, x, n, code, sr, si, stream):
        v = _arr(x, n, self._dt(code))
        v *= (complex(sr, si) if np.iscomplexobj(v) else sr)
        return 0


# ---- the kernels of the truncated sector SVD (contracts: include/gtn_b200.h) -----------------------------------
def _i64(ptr, n):
    return np.ctypeslib.as_array((C.c_int64 * n).from_address(ptr.value)) if n else np.zeros(0, np.int64)


def _i32(ptr, n):
    return np.ctypeslib.as_array((C.c_int32 * n).from_address(ptr.value)) if n else np.zeros(0, np.int32)


def _pivoted_cholesky(G, rel_thr):
    """P^T G P = L L^H stopped at numerical rank r (remaining diagonal <= rel_thr * first pivot)"""
    n = G.shape[0]
    A = G.copy()
    perm = np.arange(n)
    L = np.zeros_like(A)
    first, r = None, 0
    for j in range(n):
        d = np.real(np.diag(A))[j:]
        k = j + int(np.argmax(d))
        piv = float(np.real(A[k, k]))
        if first is None:
            first = piv
        if not (piv > rel_thr * first) or piv <= 0.0:
            break
        if k != j:
            A[[j, k], :] = A[[k, j], :]
            A[:, [j, k]] = A[:, [k, j]]
            L[[j, k], :] = L[[k, j], :]
            perm[[j, k]] = perm[[k, j]]
        L[j, j] = np.sqrt(piv)
        L[j + 1:, j] = A[j + 1:, j] / L[j, j]
        A[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j].conj())
        r = j + 1
    return L, perm, r


class HostLibSVD(HostLib):
    def _gram_sum(self, G, off, n, nsplit, dt):
        g = np.zeros((n, n), dtype=dt)
        for s_ in range(nsplit):
            g += _arr(C.c_void_p(G.value + (off + s_ * n * n) * g.itemsize), n * n, dt).reshape(n, n)
        return g

    def gtn_chol_whiten(self, G, T, code, g_off, t_off, n_dev, nprob, max_n, nsplit, rel_thr, kept, scratch, stream):
        dt = self._dt(code)
        go, to, ns, kp = _i64(g_off, nprob), _i64(t_off, nprob), _i32(n_dev, nprob), _i32(kept, nprob)
        for b in range(nprob):
            n = int(ns[b])
            g = self._gram_sum(G, int(go[b]), n, nsplit, dt)
            L, perm, r = _pivoted_cholesky(g, rel_thr)
            Tm = np.zeros((n, n), dtype=dt)
            if r:
                Pt = np.zeros((n, n))
                Pt[np.arange(n), perm] = 1.0                         # (P^T x)_i = x_perm[i]
                Tm[:r, :] = np.linalg.solve(L[:r, :r], Pt[:r, :].astype(dt))
            _arr(C.c_void_p(T.value + int(to[b]) * Tm.itemsize), n * n, dt)[:] = Tm.ravel()
            kp[b] = r
        return 0

    def gtn_gram_shift(self, G, code, g_off, n_dev, nprob, nsplit, rel_shift, stream):
        dt = self._dt(code)
        go, ns = _i64(g_off, nprob), _i32(n_dev, nprob)
        for b in range(nprob):
            n = int(ns[b])
            sl = _arr(C.c_void_p(G.value + int(go[b]) * np.dtype(dt).itemsize), nsplit * n * n, dt).reshape(nsplit, n, n)
            tr = float(np.real(np.trace(sl.sum(axis=0))))
            sl[0][np.arange(n), np.arange(n)] += rel_shift * tr
        return 0

    def gtn_gram_rotate(self, G, T, code, g_off, t_off, n_dev, nprob, max_n, nsplit, rel_thr, tol, max_sweeps,
                        sweeps, stream):
        dt = self._dt(code)
        go, to, ns = _i64(g_off, nprob), _i64(t_off, nprob), _i32(n_dev, nprob)
        for b in range(nprob):
            n = int(ns[b])
            g = self._gram_sum(G, int(go[b]), n, nsplit, dt)
            w, E_ = np.linalg.eigh((g + g.conj().T) / 2)
            Tm = E_[:, ::-1].conj().T.astype(dt)                     # unitary; rows of T B orthogonal
            _arr(C.c_void_p(T.value + int(to[b]) * Tm.itemsize), n * n, dt)[:] = np.ascontiguousarray(Tm).ravel()
        if sweeps is not None and getattr(sweeps, "value", None):
            _i32(sweeps, nprob)[:] = 1
        return 0

    def _probs(self, probs, nprob):
        from grassmanntn_b200._cabi import SvdProblem
        return (SvdProblem * nprob).from_address(probs.value)

    def gtn_jacobi_init(self, W, Z, code, probs, nprob, max_p, rn2, fro2, rn_off, stream):
        dt = self._dt(code)
        ro, f2 = _i64(rn_off, nprob), _arr(fro2, nprob, np.float64)
        for b, pr in enumerate(self._probs(probs, nprob)):
            w = _arr(C.c_void_p(W.value + pr.w_off * np.dtype(dt).itemsize), pr.p * pr.q, dt).reshape(pr.p, pr.q)
            _arr(C.c_void_p(Z.value + pr.z_off * np.dtype(dt).itemsize), pr.p * pr.p, dt)[:] = np.eye(pr.p, dtype=dt).ravel()
            n2 = np.sum(np.abs(w) ** 2, axis=1)
            _arr(C.c_void_p(rn2.value + int(ro[b]) * 8), pr.p, np.float64)[:] = n2
            f2[b] = n2.max() if pr.p else 0.0
        return 0

    def gtn_jacobi_persistent(self, W, Z, code, probs, nprob, max_p, tol, offd, rn2, fro2, rn_off, max_sweeps,
                              sweeps, stream, early_stop=0.0):
        """W0 = Z^H diag(s) Vh with the rows of W orthogonal: W <- diag(s) Vh, Z <- U^H from numpy's SVD of W0"""
        dt = self._dt(code)
        isz = np.dtype(dt).itemsize
        ro = _i64(rn_off, nprob)
        for b, pr in enumerate(self._probs(probs, nprob)):
            w = _arr(C.c_void_p(W.value + pr.w_off * isz), pr.p * pr.q, dt).reshape(pr.p, pr.q)
            z = _arr(C.c_void_p(Z.value + pr.z_off * isz), pr.p * pr.p, dt).reshape(pr.p, pr.p)
            u, s_, vh = np.linalg.svd(z.conj().T @ w, full_matrices=False)
            w[...] = s_[:, None] * vh
            z[...] = u.conj().T
            _arr(C.c_void_p(rn2.value + int(ro[b]) * 8), pr.p, np.float64)[:] = s_ ** 2
        sw = _i32(sweeps, 4)
        sw[0], sw[1] = 1, 1
        return 0

    def gtn_jacobi_finish(self, W, Z, U_out, Vh_out, s_out, code, probs, outs, order, nscratch, nprob, max_p, max_q,
                          stream):
        from grassmanntn_b200._cabi import SvdOut
        dt = self._dt(code)
        isz = np.dtype(dt).itemsize
        outs_ = (SvdOut * nprob).from_address(outs.value)
        for pr, ou in zip(self._probs(probs, nprob), outs_):
            w = _arr(C.c_void_p(W.value + pr.w_off * isz), pr.p * pr.q, dt).reshape(pr.p, pr.q).copy()
            z = _arr(C.c_void_p(Z.value + pr.z_off * isz), pr.p * pr.p, dt).reshape(pr.p, pr.p).copy()
            nrm = np.sqrt(np.sum(np.abs(w) ** 2, axis=1))
            idx = np.argsort(-nrm, kind="stable")
            _arr(C.c_void_p(s_out.value + ou.s_off * 8), pr.p, np.float64)[:] = nrm[idx]
            _arr(C.c_void_p(U_out.value + ou.u_off * isz), pr.p * pr.p, dt)[:] = z.conj().T[:, idx].ravel()
            vh = np.where(nrm[idx, None] > 0, w[idx] / np.maximum(nrm[idx, None], 1e-300), 0.0)
            _arr(C.c_void_p(Vh_out.value + pr.w_off * isz), pr.p * pr.q, dt)[:] = vh.ravel()
        return 0


class _NoStream:
    def synchronize(self):
        pass


def install(monkeypatch, truncated=False):
    """patch the product for one test; returns the package"""
    import grassmanntn_b200 as gtn
    from grassmanntn_b200 import _cabi, _engine as E, _ops
    cpu = torch.device("cpu")
    fake = (HostLibSVD if truncated else HostLib)(_cabi.lib)
    saved = dict(E._plan_cache)
    E._plan_cache.clear()
    from grassmanntn_b200 import sharded
    for name, val in (("_stream", lambda: None), ("PermutePlan", HostPlan), ("lib", fake)):
        monkeypatch.setattr(sharded, name, val)
    for mod in (E, _ops):
        monkeypatch.setattr(mod, "require_cuda", lambda: cpu)
        monkeypatch.setattr(mod, "_stream", lambda: None)
        monkeypatch.setattr(mod, "PermutePlan", HostPlan)
        monkeypatch.setattr(mod, "GemmPlan", HostGemm)
        monkeypatch.setattr(mod, "batched_svd", host_batched_svd)
        monkeypatch.setattr(mod, "lib", fake)
    monkeypatch.setattr(_cabi, "lib", fake)
    monkeypatch.setattr(_ops, "TRUNCATED_SVD", bool(truncated))
    if truncated:
        # the subspace iteration with its certificate, eager launches only (no CUDA graphs, no speculation)
        monkeypatch.setattr(E, "USE_GRAPHS", False)
        monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _NoStream())
        monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
        for name in ("_trunc_plans", "_trunc_iters_hint", "_trunc_rate", "_trunc_fail", "_trunc_probe"):
            monkeypatch.setattr(E, name, {})
    monkeypatch.setattr(gtn.gauge2d, "SPECULATE", False)
    monkeypatch.setattr(gtn.gauge2d, "STEP_GRAPH", False)
    return gtn, saved


def uninstall(saved):
    from grassmanntn_b200 import _engine as E
    E._plan_cache.clear()
    E._plan_cache.update(saved)
