"""Index/sign primitives (host side).

Mirrors the reference module `param` (param.py:61-82): `gparity`, `sgn`, `encoder` with the
same names and integer results, computed with bit tricks instead of three 65536-entry literal
tables (param.py:4-6) and valid for any non-negative int.  On the device the same quantities
are derived per element from popcounts inside the sign+permute kernel
(csrc/gtn_permute.cu); these host versions are used by the planner that builds the kernel's
parity tables and by user code.
"""
import numpy as np


def gparity(i):
    """Grassmann parity count of index i = popcount(i) (param.py:61-66)."""
    return int(i).bit_count()


def sgn(i):
    """sigma_i = (-1)^(p(p-1)/2) with p = popcount(i) (param.py:68-73)."""
    return -1 if (int(i).bit_count() >> 1) & 1 else 1


def encoder(i):
    """canonical <-> parity-preserving index map, self-inverse (param.py:75-82)."""
    i = int(i)
    return i ^ ((i >> 1).bit_count() & 1)


def _popcount_vec(x):
    x = np.asarray(x, dtype=np.uint64).copy()
    c = np.zeros(x.shape, dtype=np.int64)
    while np.any(x):
        c += (x & np.uint64(1)).astype(np.int64)
        x >>= np.uint64(1)
    return c


def popcount_array(n):
    """popcount(0..n-1) as int64 array."""
    return _popcount_vec(np.arange(n, dtype=np.uint64))


def encoder_array(n):
    i = np.arange(n, dtype=np.int64)
    return i ^ (_popcount_vec(i >> 1) & 1)


def canonical_of_block(pi, s):
    """canonical index of block element (parity pi, offset s): encoder(2s+pi)
    (reference block.__init__ __init__.py:300-303)."""
    s = np.asarray(s, dtype=np.int64)
    return 2 * s + (pi ^ (_popcount_vec(s) & 1))
