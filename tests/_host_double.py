"""TEST DOUBLE for the CUDA library (tests only; nothing under grassmanntn_b200/ imports this).

`install(monkeypatch)` lets the product's HOST logic -- planner, launch tables, GEMM group lists, decomposition
packing, rank rule, unpacking -- run without a GPU by standing in for the C ABI with numpy:
  gtn_sign_permute      -> the numpy emulation of the kernel's addressing / sign rule (test_tables_cpu.emulate)
  gtn_grouped_gemm      -> numpy matmul per group (offsets, leading dimensions, batch strides, alpha/beta, conj-trans B)
  gtn_jacobi_* (full)   -> numpy.linalg.svd behind _engine.batched_svd's interface
  gtn_sumsq / rowsum / row_sumsq / dot / pow_rcond / scale -> numpy on the host buffers
The truncated-SVD path, CUDA graphs and the multi-GPU mode are NOT emulated (GPU tests cover them).
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


class HostPlan:
    def __init__(self, jobs):
        self.jobs = jobs

    def run(self, src, dst, scale=1.0):
        s, d = src.numpy(), dst.numpy()
        for f, tabs in self.jobs:
            emulate(f, tabs, s, d, scale)


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

    def gtn_rowsum(self, x, y, rows, cols, code, stream):
        dt = self._dt(code)
        _arr(y, rows, dt)[:] = _arr(x, rows * cols, dt).reshape(rows, cols).sum(axis=1)
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


def install(monkeypatch):
    """patch the product for one test; returns the package"""
    import grassmanntn_b200 as gtn
    from grassmanntn_b200 import _cabi, _engine as E, _ops
    cpu = torch.device("cpu")
    fake = HostLib(_cabi.lib)
    saved = dict(E._plan_cache)
    E._plan_cache.clear()
    for mod in (E, _ops):
        monkeypatch.setattr(mod, "require_cuda", lambda: cpu)
        monkeypatch.setattr(mod, "_stream", lambda: None)
        monkeypatch.setattr(mod, "PermutePlan", HostPlan)
        monkeypatch.setattr(mod, "GemmPlan", HostGemm)
        monkeypatch.setattr(mod, "batched_svd", host_batched_svd)
        monkeypatch.setattr(mod, "lib", fake)
    monkeypatch.setattr(_cabi, "lib", fake)
    monkeypatch.setattr(_ops, "TRUNCATED_SVD", False)
    monkeypatch.setattr(gtn.gauge2d, "SPECULATE", False)
    monkeypatch.setattr(gtn.gauge2d, "STEP_GRAPH", False)
    return gtn, saved


def uninstall(saved):
    from grassmanntn_b200 import _engine as E
    E._plan_cache.clear()
    E._plan_cache.update(saved)
