"""Generate the golden vectors under tests/golden/ from the REAL reference (build container only).

    python tests/golden/make_golden.py

Imports the unmodified reference through oracle/ref_harness.py (stand-ins for the missing
`sparse` / `opt_einsum` packages, arith.exp alias, fixed power_block -- all documented there) and
stores, for seeded inputs (numpy global RNG, the RNG gtn.random uses, reference __init__.py:6036):
  param_tables.npz      the three 65536-entry tables of param.py:4-6 (gparity / encoder / sgn)
  einsum_cases.npz      inputs + outputs of a sweep of gtn.einsum strings (dense)
  decomp_cases.npz      singular values / eigenvalues, reconstructions, hconjugate outputs
  cg_random.npz         per-step Tnorm / trace error / logZ of trg, atrg2dy, atrg2dx on a random
                        Grassmann-even tensor, dense and block formats
The Z2 fixtures are made by make_z2_tensor.py / make_z2_cg.py.
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_harness

gtn = ref_harness.load_reference()
from grassmanntn import param as rparam

# ---------------------------------------------------------------- param tables
np.savez_compressed(os.path.join(HERE, "param_tables.npz"),
                    gparity=np.asarray(rparam.arr_gparity), encoder=np.asarray(rparam.arr_encoder),
                    sgn=np.asarray(rparam.arr_sgn),
                    big=np.array([[i, rparam.gparity(i), rparam.sgn(i), rparam.encoder(i)]
                                  for i in (65536, 65537, 100000, 131071, 262145)]))

# ---------------------------------------------------------------- einsum sweep
EINSUM_CASES = [
    ('ijkl->jkli', [((4, 4, 4, 4), (1, 1, -1, -1))]),
    ('ijkl->klij', [((4, 2, 8, 4), (1, -1, -1, 1))]),
    ('ijkl->lijk', [((8, 4, 8, 4), (1, 1, -1, -1))]),
    ('ijkl,klmn->ijmn', [((4, 4, 4, 2), (1, 1, 1, 1)), ((4, 2, 4, 8), (-1, -1, 1, -1))]),
    ('kwz,lxw->lxzk', [((4, 4, 8), (1, -1, 1)), ((4, 2, 4), (-1, 1, 1))]),
    ('yxi,zyj->jzxi', [((4, 2, 8), (1, -1, 1)), ((4, 4, 2), (1, -1, -1))]),
    ('lxzk,jzxi->ijkl', [((4, 2, 8, 4), (1, 1, -1, 1)), ((2, 8, 2, 4), (-1, 1, -1, 1))]),
    ('ijij', [((4, 2, 4, 2), (1, 1, -1, -1))]),
    ('ijij', [((4, 2, 4, 2), (-1, 1, 1, -1))]),
    ('ijkl,klij', [((4, 2, 4, 8), (1, 1, -1, -1)), ((4, 8, 4, 2), (1, 1, -1, -1))]),
    ('IJIK,iKiJ', [((4, 2, 4, 8), (1, 1, -1, -1)), ((2, 8, 2, 2), (1, 1, -1, -1))]),
    ('IJKLij,ij->IJKL', [((4, 2, 4, 2, 3, 3), (1, 1, -1, -1, 0, 0)), ((3, 3), (0, 0))]),
    ('i1 i3 a, j1 j3 b -> i1 i3 ab j1 j3', [((4, 2, 3), (1, -1, 0)), ((2, 4, 2), (1, -1, 0))]),
    ('t s al m , lbm -> t s ab m', [((4, 2, 4, 2, 3), (1, -1, 1, 1, 0)), ((2, 4, 3), (-1, 1, 0))]),
    ('abx,xc->abc', [((4, 4, 8), (1, 1, 1)), ((8, 8), (-1, 1))]),
    ('ax,xbc->abc', [((8, 8), (-1, 1)), ((8, 4, 4), (-1, -1, -1))]),
    ('ajk,jib->aibk', [((4, 2, 8), (-1, -1, -1)), ((2, 4, 4), (1, 1, 1))]),
    ('iax,xbj->ijab', [((4, 2, 8), (1, 1, 1)), ((8, 4, 2), (-1, 1, -1))]),
    ('ab,bc,cd->ad', [((4, 8), (1, 1)), ((8, 2), (-1, 1)), ((2, 4), (-1, -1))]),
    ('abc,dbe,fce->adf', [((4, 8, 2), (1, 1, 1)), ((2, 8, 4), (1, -1, 1)), ((4, 2, 4), (-1, -1, -1))]),
    ('IJIJmn,KLKLmn->mn', [((4, 2, 4, 2, 3, 2), (1, 1, -1, -1, 0, 0)), ((2, 2, 2, 2, 3, 2), (1, 1, -1, -1, 0, 0))]),
    ('xa,xb->ab', [((4, 2), (1, 1)), ((4, 8), (-1, -1))]),
]
out = {"n": len(EINSUM_CASES)}
for k, (sub, ops) in enumerate(EINSUM_CASES):
    np.random.seed(100 + k)
    objs = [gtn.random(s, st, dtype=complex, skip_trimming=(k % 2 == 0)) for s, st in ops]
    res = gtn.einsum(sub, *objs)
    out["sub_%d" % k] = sub
    out["nops_%d" % k] = len(ops)
    for j, (o, (s, st)) in enumerate(zip(objs, ops)):
        out["in_%d_%d" % (k, j)] = o.data
        out["st_%d_%d" % (k, j)] = np.array(st)
    if np.ndim(res) == 0 or np.isscalar(res):
        out["out_%d" % k] = np.array(res)
        out["ost_%d" % k] = np.array([])
    else:
        out["out_%d" % k] = res.data
        out["ost_%d" % k] = np.array(res.statistics)
np.savez_compressed(os.path.join(HERE, "einsum_cases.npz"), **out)
print("einsum cases done", flush=True)

# ---------------------------------------------------------------- decompositions / hconjugate / switches
out = {}
np.random.seed(7)
A = gtn.random((4, 4, 4, 4), (1, 1, -1, -1), dtype=complex)
out["A"] = A.data
for cut in (None, 8, 6):
    U, S, V = A.svd('ab|cd', cut)
    tag = "svd_%s" % cut
    out[tag + "_S"] = S.data
    out[tag + "_rec"] = gtn.einsum('abx,xy,ycd->abcd', U, S, V).data
    Ub, Sb, Vb = A.toblock().svd('ab|cd', cut)
    out[tag + "_blk_even"] = np.array(Sb.even_shape)
    out[tag + "_blk_odd"] = np.array(Sb.odd_shape)
    out[tag + "_blk_Snorm"] = Sb.norm
    out[tag + "_blk_rec"] = gtn.einsum('abx,xy,ycd->abcd', Ub, Sb, Vb).todense().data
cA = A.hconjugate('ab|cd')
out["hconj_A"] = cA.data
M = gtn.einsum('abcd,cdef->abef', cA, A)
out["gram"] = M.data
U, S, V = M.eig('ab|cd', 8)
out["eig_S"] = S.data
out["eig_rec"] = gtn.einsum('abx,xy,ycd->abcd', U, S, V).data
np.random.seed(8)
Bh = gtn.random((4, 2, 3, 4, 2), (1, -1, 0, -1, 1), dtype=complex)
out["B"] = Bh.data
U, S, V = Bh.svd('abc|de')
out["svdB_S"] = S.data
out["svdB_Ustat"] = np.array([str(s) for s in U.statistics])
out["hconj_B"] = Bh.hconjugate('abc|de').data
np.random.seed(9)
Cx = gtn.random((4, 4, 2, 8), (1, -1, -1, 1), dtype=complex, skip_trimming=True)
out["C"] = Cx.data
out["C_fmt"] = Cx.switch_format().data
out["C_enc"] = Cx.switch_encoder().data
out["C_par"] = Cx.switch_parity().data
out["C_fmt_enc"] = Cx.switch_format().switch_encoder().data
out["C_sqrt_in"] = np.diag(np.array([4.0, 1e-11, 9.0, 0.25]))
Sq = gtn.sqrt(gtn.dense(np.diag(np.array([4.0, 1e-11, 9.0, 0.25])), statistics=(-1, 1)))
out["C_sqrt"] = Sq.data
np.savez_compressed(os.path.join(HERE, "decomp_cases.npz"), **out)
print("decomp cases done", flush=True)

# ---------------------------------------------------------------- CG steps on a random even tensor
out = {}
np.random.seed(21)
T0 = gtn.random((4, 4, 4, 4), (1, 1, -1, -1), dtype=complex)
out["T0"] = T0.data
for fmt, mod, cut in (("dense", gtn.gauge2d, 8), ("block", gtn.gauge2d_block, 6)):
    T = T0.copy() if fmt == "dense" else T0.toblock()
    for algo in ("trg", "atrg2dy", "atrg2dx"):
        fn = getattr(mod, algo)
        X = T.copy()
        rec = []
        for step in range(2):
            if algo == "trg":
                X, Tn, err = fn(X, cut, error_test=True)
            else:
                X, Tn, err = fn(X, X, cut, error_test=True)
            F = mod.logZ(X.copy(), 'anti-periodic')
            rec.append([Tn, err, F.real, F.imag])
        out["%s_%s" % (fmt, algo)] = np.array(rec)
        out["%s_%s_shape" % (fmt, algo)] = np.array(X.effective_shape if fmt == "block" else X.shape)
np.savez_compressed(os.path.join(HERE, "cg_random.npz"), **out)
print("cg cases done", flush=True)
