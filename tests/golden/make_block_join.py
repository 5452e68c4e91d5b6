"""Golden vectors for join_legs_block / split_legs_block (reference __init__.py:3322-3859) from the REAL
reference (build container only):  python tests/golden/make_block_join.py  -> block_join.npz

Every case: a block tensor with seeded random entries in ALL parity blocks (not trimmed), possibly ragged
even / odd dimensions, standard or matrix format; joined with a grouping string and final statistics, then
split back.  Stored: input blocks, joined blocks + the joined tensor's sgn vectors, the blocks of the split
result (must equal the input), and the joined tensor after switch_format (custom sigma vectors in use)."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_harness

gtn = ref_harness.load_reference()

# (even_shape, odd_shape, statistics, format, grouping string, final_stat)
CASES = [
    ((2, 2, 2, 2), (2, 2, 2, 2), (1, 1, -1, -1), "standard", "(ij)(kl)", (1, -1)),
    ((2, 2, 2, 2), (2, 2, 2, 2), (1, 1, -1, -1), "standard", "(ij)(kl)", (-1, 1)),
    ((2, 1, 4, 2), (2, 1, 4, 2), (1, -1, -1, 1), "matrix", "(ij)(kl)", (-1, -1)),
    ((3, 2, 2, 1), (2, 3, 1, 2), (1, 1, -1, -1), "standard", "(ij)(kl)", (1, -1)),
    ((3, 2, 2, 1), (2, 3, 1, 2), (1, -1, 1, -1), "matrix", "(ijk)(l)", (-1, 1)),
    ((2, 3, 2), (1, 2, 2), (1, 1, -1), "standard", "(ijk)", (-1,)),
    ((2, 2, 3, 2), (2, 1, 3, 2), (1, -1, 0, 0), "standard", "(ij)(kl)", (1, 0)),
    ((2, 3, 2, 2, 2), (2, 3, 1, 2, 3), (-1, 0, 1, 1, -1), "standard", "(i)(j)(klm)", (-1, 0, -1)),
    ((2, 2, 1, 2, 2), (1, 2, 2, 2, 1), (1, 1, 1, -1, -1), "matrix", "(ijk)(lm)", (1, -1)),
    ((2, 2, 2, 2, 2, 2), (2, 2, 2, 2, 2, 2), (1, 1, 1, 1, -1, -1), "standard", "(ijkl)(mn)", (-1, 1)),
]

out = {"n": len(CASES)}
for k, (ev, od, st, fmt, string, fstat) in enumerate(CASES):
    np.random.seed(500 + k)
    B = gtn.zero_block_eo(ev, od, st, format=fmt, dtype=complex)
    it = np.nditer(B.data, flags=["multi_index", "refs_ok"])
    pats = []
    for _ in it:
        p = it.multi_index
        shp = B.data[p].shape
        B.data[p] = np.random.rand(*shp) + 1j * np.random.rand(*shp)
        pats.append(p)
        out["c%d_in_%s" % (k, "".join(map(str, p)))] = B.data[p].copy()
    J = gtn.join_legs_block(B, string, fstat)
    assert J.marked_as_joined
    it = np.nditer(J.data, flags=["multi_index", "refs_ok"])
    for _ in it:
        p = it.multi_index
        out["c%d_join_%s" % (k, "".join(map(str, p)))] = np.asarray(J.data[p]).copy()
    for pi in (0, 1):
        for ax in range(J.ndim):
            out["c%d_sgn_%d_%d" % (k, pi, ax)] = np.asarray(J.sgn[pi][ax]).astype(np.int64)
    out["c%d_join_even" % k] = np.asarray(J.even_shape, dtype=np.int64)
    out["c%d_join_odd" % k] = np.asarray(J.odd_shape, dtype=np.int64)
    out["c%d_join_shape" % k] = np.asarray(J.shape, dtype=np.int64)
    out["c%d_join_format" % k] = np.str_(J.format)
    Jm = J.switch_format()
    it = np.nditer(Jm.data, flags=["multi_index", "refs_ok"])
    for _ in it:
        p = it.multi_index
        out["c%d_joinsw_%s" % (k, "".join(map(str, p)))] = np.asarray(Jm.data[p]).copy()
    S = gtn.split_legs_block(J, string, st, B.shape, ev, od)
    worst = 0.0
    for p in pats:
        out["c%d_split_%s" % (k, "".join(map(str, p)))] = np.asarray(S.data[p]).copy()
        worst = max(worst, float(np.abs(S.data[p] - B.data[p]).max()) if B.data[p].size else 0.0)
    out["c%d_split_format" % k] = np.str_(S.format)
    out["c%d_shape" % k] = np.asarray(B.shape, dtype=np.int64)
    print(k, st, fmt, string, fstat, "joined even/odd", J.even_shape, J.odd_shape, "round trip", worst, flush=True)
np.savez_compressed(os.path.join(HERE, "block_join.npz"), **out)
print("saved block_join.npz")
