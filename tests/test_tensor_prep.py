"""Initial-tensor compression pipeline (reference gauge2d.py:22-66, :1198-1585):
grassmanntn_b200.gauge2d.fcompress_B / compress_B / compress_A / compress_T / tensor_from_AB against the REAL reference's
functions on the Z2 model of example.py's defaults (tests/golden/make_z2_prep_golden.py).  The isometries carry an SVD
gauge, so the stages are compared through gauge-invariant numbers: shapes (the bond dimensions the rank rule and the
cut produce), norms of the compressed tensors, z1, z4 and the trace error -- all to 1e-10 relative -- and the final site
tensor through one TRG step against the committed fixture tensor (Tnorm and F to 1e-10)."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _inputs(gtn):
    z = np.load(os.path.join(G, "z2_prep.npz"))
    B = np.zeros(tuple(int(x) for x in z["B_shape"]), dtype=complex)
    B[tuple(z["B_coords"].T.astype(np.int64))] = z["B_vals"]
    A = gtn.dense(np.asarray(z["A"]).astype(complex), statistics=tuple(int(s) for s in z["A_stats"]))
    return z, A, gtn.dense(B, statistics=tuple(int(s) for s in z["B_stats"]))


def run_pipeline(gtn):
    g = gtn.gauge2d
    E = gtn.einsum
    z, A, B = _inputs(gtn)
    rel = lambda a, b: abs(a - b) / abs(b)
    z1 = E("IJIJijij,jiji", B, A)
    assert rel(z1, complex(z["z1"])) <= 1e-12
    B = g.fcompress_B(B)
    assert tuple(B.shape) == tuple(z["fB_shape"]) and rel(B.norm, float(z["fB_norm"])) <= 1e-10
    B, Us = g.compress_B(B)
    assert tuple(B.shape) == tuple(z["cB_shape"]) and rel(B.norm, float(z["cB_norm"])) <= 1e-10
    assert [tuple(u.shape) for u in Us] == [tuple(s) for s in z["U_shapes"]]
    A = g.compress_A(A, Us)
    assert tuple(A.shape) == tuple(z["cA_shape"]) and rel(A.norm, float(z["cA_norm"])) <= 1e-10
    T = E('IJXYijklmn,XYKL->IJKLijklmn', A, B)
    assert tuple(T.shape) == tuple(z["T0_shape"]) and rel(T.norm, float(z["T0_norm"])) <= 1e-10
    T = g.compress_T(T)
    assert tuple(T.shape) == tuple(z["T_shape"]) and rel(T.norm, float(z["T_norm"])) <= 1e-10
    assert tuple(T.statistics) == (1, 1, -1, -1, 0, 0)
    z4 = E("IJIJij,ij", T, gtn.dense(np.full((2, 2), 1.0), statistics=(0, 0)))
    assert rel(z4, complex(z["z4"])) <= 1e-10 and abs(1 - z4 / z1) <= 1e-12
    # the whole pipeline in one call, and its product in a coarse-graining step next to the reference's fixture tensor
    _, A0, B0 = _inputs(gtn)
    T2, err = g.tensor_from_AB(A0, B0)
    assert err <= 1e-12 and rel(T2.norm, float(z["T_norm"])) <= 1e-10
    Tf = g.load_initial_tensor()
    a, na = g.trg(g.zcap(T2), 16)
    b, nb = g.trg(g.zcap(Tf), 16)
    assert rel(na, nb) <= 1e-10
    assert rel(g.logZ(a, "anti-periodic"), g.logZ(b, "anti-periodic")) <= 1e-10


@pytest.mark.gpu
def test_gpu_tensor_preparation_vs_reference(gtn):
    run_pipeline(gtn)


@pytest.mark.gpu
def test_gpu_bosonic_diagonal_einsum(gtn):
    """bosonic indices repeated inside an operand (its diagonal), as in 'IJIJijij,jiji' (reference gauge2d.py:32)"""
    import gtn_oracle as O
    rng = np.random.RandomState(4)
    for sub, shapes, stats in [("IJIJijij,jiji", [(4, 4, 4, 4, 2, 3, 2, 3), (3, 2, 3, 2)], [(1, 1, -1, -1, 0, 0, 0, 0), (0, 0, 0, 0)]),
                               ("iij,jk->ik", [(2, 2, 3), (3, 2)], [(0, 0, 0), (0, 0)]),
                               ("IJKLijkl,km->JKLjklmIi", [(4, 4, 4, 4, 2, 2, 2, 2), (2, 2)], [(1, 1, -1, -1, 0, 0, 0, 0), (0, 0)])]:
        os_ = [O.random_dense(sh, st, dtype=complex, rng=rng) for sh, st in zip(shapes, stats)]
        gs = [gtn.dense(o.data, statistics=o.statistics) for o in os_]
        r, ro = gtn.einsum(sub, *gs), O.einsum(sub, *os_)
        if isinstance(ro, O.Dense):
            assert np.abs(np.asarray(r.data.cpu()) - ro.data).max() <= 1e-12 * np.abs(ro.data).max(), sub
        else:
            assert abs(r - ro) <= 1e-12 * abs(ro), sub
