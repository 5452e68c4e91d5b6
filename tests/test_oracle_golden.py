"""CPU: the oracle (oracle/gtn_oracle.py) against golden vectors produced by the REAL reference
(tests/golden/make_golden.py) and against the reference's own literal tables / doc numbers."""
import os

import numpy as np
import pytest

import gtn_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def test_param_tables_bit_exact():
    z = _load("param_tables.npz")
    n = len(z["gparity"])
    assert n == 65536
    assert np.array_equal(z["gparity"], np.array([O.gparity(i) for i in range(n)]))
    assert np.array_equal(z["encoder"], np.array([O.encoder(i) for i in range(n)]))
    assert np.array_equal(z["sgn"], np.array([O.sgn(i) for i in range(n)]))
    for i, gp, sg, en in z["big"]:
        assert (O.gparity(i), O.sgn(i), O.encoder(i)) == (gp, sg, en)


def test_product_param_matches_tables():
    from grassmanntn_b200 import param
    z = _load("param_tables.npz")
    n = 65536
    assert np.array_equal(z["gparity"], param.popcount_array(n))
    assert np.array_equal(z["encoder"], param.encoder_array(n))
    assert all(param.sgn(i) == z["sgn"][i] for i in range(0, n, 7))
    assert all(param.encoder(param.encoder(i)) == i for i in range(0, 1 << 20, 4099))


def test_einsum_cases():
    z = _load("einsum_cases.npz")
    for k in range(int(z["n"])):
        sub = str(z["sub_%d" % k])
        ops = [O.Dense(z["in_%d_%d" % (k, j)], [int(s) for s in z["st_%d_%d" % (k, j)]])
               for j in range(int(z["nops_%d" % k]))]
        res = O.einsum(sub, *ops)
        ref = z["out_%d" % k]
        if isinstance(res, O.Dense):
            assert tuple(res.statistics) == tuple(int(s) for s in z["ost_%d" % k])
            assert np.abs(res.data - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1), sub
        else:
            assert abs(res - ref) <= 1e-13 * max(abs(ref), 1), sub


def test_decompositions_and_switches():
    z = _load("decomp_cases.npz")
    A = O.Dense(z["A"], (1, 1, -1, -1))
    for cut in (None, 8, 6):
        tag = "svd_%s" % cut
        U, S, V = O.svd(A, 'ab|cd', cut)
        assert np.abs(S.data - z[tag + "_S"]).max() <= 1e-12 * np.abs(z[tag + "_S"]).max()
        rec = O.einsum('abx,xy,ycd->abcd', U, S, V).data
        assert np.abs(rec - z[tag + "_rec"]).max() <= 1e-12 * np.abs(rec).max()
        Ub, Sb, Vb, (nE, nO) = O.svd(A, 'ab|cd', cut, rule="block", return_counts=True)
        d = max(nE, nO)
        assert tuple(z[tag + "_blk_even"]) == (d, d) and tuple(z[tag + "_blk_odd"]) == (d, d)
        assert abs(Sb.norm - float(z[tag + "_blk_Snorm"])) <= 1e-12 * Sb.norm
        recb = O.einsum('abx,xy,ycd->abcd', Ub, Sb, Vb).data
        assert np.abs(recb - z[tag + "_blk_rec"]).max() <= 1e-12 * np.abs(recb).max()
    cA = O.hconjugate(A, 'ab|cd')
    assert np.array_equal(cA.data, z["hconj_A"])
    M = O.einsum('abcd,cdef->abef', cA, A)
    assert np.abs(M.data - z["gram"]).max() <= 1e-12 * np.abs(M.data).max()
    U, S, V = O.eig(O.Dense(z["gram"], M.statistics), 'ab|cd', 8)
    assert np.abs(S.data - z["eig_S"]).max() <= 1e-10 * np.abs(z["eig_S"]).max()
    B = O.Dense(z["B"], (1, -1, 0, -1, 1))
    U, S, V = O.svd(B, 'abc|de')
    assert np.abs(S.data - z["svdB_S"]).max() <= 1e-12 * np.abs(S.data).max()
    assert [str(s) for s in U.statistics] == [str(s) for s in z["svdB_Ustat"]]
    assert np.array_equal(O.hconjugate(B, 'abc|de').data, z["hconj_B"])
    C = O.Dense(z["C"], (1, -1, -1, 1))
    assert np.array_equal(C.switch_format().data, z["C_fmt"])
    assert np.array_equal(C.switch_encoder().data, z["C_enc"])
    assert np.array_equal(C.switch_parity().data, z["C_par"])
    assert np.array_equal(C.switch_format().switch_encoder().data, z["C_fmt_enc"])
    sq = O.sqrt(O.Dense(z["C_sqrt_in"], (-1, 1)))
    assert np.allclose(sq.data, z["C_sqrt"], rtol=1e-15, atol=0, equal_nan=True)


@pytest.mark.parametrize("fmt,cut", [("dense", 8), ("block", 6)])
@pytest.mark.parametrize("algo", ["trg", "atrg2dy", "atrg2dx"])
def test_cg_steps(algo, fmt, cut):
    z = _load("cg_random.npz")
    X = O.Dense(z["T0"], (1, 1, -1, -1))
    ref = z["%s_%s" % (fmt, algo)]
    for step in range(2):
        if algo == "trg":
            X, Tn, err = O.trg(X, cut, rule=fmt, error_test=True)
        else:
            X, Tn, err = getattr(O, algo)(X, X, cut, rule=fmt, error_test=True)
        F = O.logZ(X, 'anti-periodic', block_format=(fmt == "block"))
        assert abs(Tn - ref[step, 0]) <= 1e-10 * ref[step, 0]
        assert abs(err - ref[step, 1]) <= 1e-8 * max(ref[step, 1], 1e-3)
        assert abs(F - complex(ref[step, 2], ref[step, 3])) <= 1e-10 * max(abs(F), 1)
