"""ctypes binding of libgtn_b200.so (C ABI declared in include/gtn_b200.h).

There is NO CPU fallback: importing this module without the built library raises, and every
wrapper raises on a non-zero status.  The library is built in-tree by
`__graft_entry__.build()` / `make -C grassmanntn_b200/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgtn_b200.so")

GTN_F64, GTN_C128 = 0, 1
GTN_MAX_SUPER = 8


class AxisEntry(C.Structure):
    _fields_ = [("in_off", C.c_int64), ("out_off", C.c_int64), ("P", C.c_uint32), ("M", C.c_uint32)]


class PermuteJob(C.Structure):
    _fields_ = [("in_base", C.c_int64), ("out_base", C.c_int64),
                ("table_start", C.c_int64 * GTN_MAX_SUPER), ("size", C.c_int32 * GTN_MAX_SUPER),
                ("nsuper", C.c_int32), ("const_exp", C.c_int32), ("conj", C.c_int32),
                ("transpose", C.c_int32), ("ntiles", C.c_int64)]


class GemmGroup(C.Structure):
    _fields_ = [("a_off", C.c_int64), ("b_off", C.c_int64), ("c_off", C.c_int64),
                ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64),
                ("batch_stride_a", C.c_int64), ("batch_stride_b", C.c_int64), ("batch_stride_c", C.c_int64),
                ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32), ("batch", C.c_int32),
                ("alpha", C.c_double), ("beta", C.c_double), ("tile_start", C.c_int64),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class SvdProblem(C.Structure):
    _fields_ = [("w_off", C.c_int64), ("z_off", C.c_int64), ("p", C.c_int32), ("q", C.c_int32)]


class SvdOut(C.Structure):
    _fields_ = [("s_off", C.c_int64), ("u_off", C.c_int64)]


class SvdInfo(C.Structure):
    _fields_ = [("start_iters", C.c_int32), ("iters", C.c_int32), ("checks", C.c_int32), ("sweeps", C.c_int32),
                ("launches", C.c_int32), ("robust", C.c_int32), ("worst", C.c_double), ("rate", C.c_double)]


GTN_OP_SECTOR_SVD_TRUNC, GTN_OP_SECTOR_EIGH_TRUNC = 0, 1
GTN_ERR_NOT_CONVERGED = -3


class GtnError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "grassmanntn_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C grassmanntn_b200/csrc` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    sigs = {
        "gtn_sign_permute": (i32, [vp, vp, i32, vp, vp, i32, i64, dbl, dbl, vp]),
        "gtn_gemm_plan_host": (i64, [C.POINTER(GemmGroup), i32, i32, i32]),
        "gtn_grouped_gemm": (i32, [vp, vp, vp, i32, vp, i32, i64, i32, vp]),
        "gtn_gemm_tma_check": (i32, [C.POINTER(GemmGroup), i32, i32]),
        "gtn_grouped_gemm_tma": (i32, [vp, vp, vp, i32, C.POINTER(GemmGroup), vp, i32, i64, i32, vp]),
        "gtn_grouped_gemm_bcast": (i32, [vp, vp, vp, i32, i32, vp, i32, i64, vp]),
        "gtn_jacobi_init": (i32, [vp, vp, i32, vp, i32, i32, vp, vp, vp, vp]),
        "gtn_jacobi_sweep": (i32, [vp, vp, i32, vp, i32, i32, i32, dbl, vp, vp, vp, vp, vp]),
        "gtn_jacobi_persistent": (i32, [vp, vp, i32, vp, i32, i32, dbl, vp, vp, vp, vp, i32, vp, vp, dbl]),
        "gtn_jacobi_finish": (i32, [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, vp]),
        "gtn_small_eigh_whiten": (i32, [vp, vp, i32, vp, vp, vp, i32, i32, dbl, vp, vp, vp, vp]),
        "gtn_chol_whiten": (i32, [vp, vp, i32, vp, vp, vp, i32, i32, i32, dbl, vp, vp, vp]),
        "gtn_chol_whiten_scratch_elems": (i64, [i32]),
        "gtn_gram_shift": (i32, [vp, i32, vp, vp, i32, i32, dbl, vp]),
        "gtn_debug_phase_clocks": (i32, [vp]),
        "gtn_gram_rotate": (i32, [vp, vp, i32, vp, vp, vp, i32, i32, i32, dbl, dbl, i32, vp, vp]),
        "gtn_sumsq": (i32, [vp, i64, i32, vp, i32, vp]),
        "gtn_sumabs": (i32, [vp, i64, i32, vp, i32, vp]),
        "gtn_rowsum": (i32, [vp, vp, i64, i64, i32, vp]),
        "gtn_dot": (i32, [vp, vp, i64, i32, vp, vp, i32, vp]),
        "gtn_row_sumsq": (i32, [vp, vp, i64, i64, i32, vp]),
        "gtn_sum_slices": (i32, [vp, vp, i64, i32, i32, vp]),
        "gtn_pow_rcond": (i32, [vp, i64, i32, dbl, dbl, vp]),
        "gtn_scale": (i32, [vp, i64, i32, dbl, dbl, vp]),
        "gtn_odd_checker": (i32, [vp, i64, i64, i32, vp, vp]),
        "gtn_workspace_bytes": (i64, [i32, i32, i32, vp, vp, vp]),
        "gtn_sector_svd_trunc": (i32, [vp, vp, vp, i32, i32, vp, dbl, vp, vp, vp, vp, vp, i64, vp, vp]),
        "gtn_sector_eigh_trunc": (i32, [vp, vp, vp, i32, i32, vp, dbl, vp, vp, vp, vp, vp, vp, i64, vp, vp]),
        "gtn_comm_available": (i32, []),
        "gtn_comm_unique_id": (i32, [vp]),
        "gtn_comm_init": (i32, [vp, i32, i32, vp]),
        "gtn_comm_destroy": (i32, [vp]),
        "gtn_allreduce": (i32, [vp, vp, i64, i32, i32, vp]),
        "gtn_allgather": (i32, [vp, vp, vp, i64, i32, vp]),
        "gtn_broadcast": (i32, [vp, vp, i64, i32, i32, vp]),
        "gtn_peer_barrier": (i32, [vp, i32, i32, C.c_uint64, vp]),
        "gtn_version": (i32, []),
        "gtn_build_arch": (C.c_char_p, []),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib, tuple(sigs)


lib, EXPORTED = _load()

# number of kernel launches issued through this binding (bench.py reports it as gpu_launches)
launch_count = 0


def check(status, what):
    if status != 0:
        raise GtnError("%s failed with status %d" % (what, status))


def count(n=1):
    global launch_count
    launch_count += n
