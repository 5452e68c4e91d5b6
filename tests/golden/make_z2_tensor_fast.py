"""Same as make_z2_tensor.py, but the reference's einsum_ds (whose Python sign-tensor loops,
__init__.py:1962-1999 / :2088-2126, make the compression stage take hours on the dense-backed
`sparse` stand-in) is replaced by the oracle's vectorised restatement oracle.gtn_oracle.einsum,
which tests/test_oracle_golden.py pins against the real einsum_ds (einsum_cases.npz).  All other
steps (fcompress_B, compress_B, compress_A, compress_T, svd/eig/hconjugate) are the reference's
own code.  Writes tests/golden/z2_initial_tensor.npz with a `generator` field saying which script
produced it; if the slow all-reference script finishes too, compare with compare_z2.py."""
import os, sys, time, pickle
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_harness
import gtn_oracle as O

gtn = ref_harness.load_reference()
g = gtn.gauge2d


def fast_einsum_ds(*args, format="standard", encoder="canonical", ignore_anticommutation=False, debug_mode=False):
    sub = args[0]
    n = len(sub.replace(" ", "").split("->")[0].split(","))
    objs = args[1:1 + n]
    this_type = type(objs[0])
    conv = []
    for o in objs:
        d = gtn.dense(o) if this_type is gtn.sparse else o
        conv.append(O.Dense(np.asarray(d.data), d.statistics, d.encoder, d.format))
    res = O.einsum(sub, *conv, ignore_anticommutation=ignore_anticommutation)
    if not isinstance(res, O.Dense):
        return res
    out = gtn.dense(res.data, statistics=res.statistics, encoder=res.encoder, format=res.format)
    return gtn.sparse(out) if this_type is gtn.sparse else out


gtn.einsum_ds = fast_einsum_ds
Nphi = 2
cache = "/tmp/z2_AB.pkl"
t0 = time.time()
Ad, Ast, Bd, Bst = pickle.load(open(cache, "rb"))
A = gtn.sparse(Ad, statistics=Ast)
B = gtn.sparse(Bd, statistics=Bst)
normA, normB, nnzB = float(A.norm), float(B.norm), int(B.nnz)
z1 = gtn.einsum("IJIJijij,jiji", B, A)
print("z1", z1, flush=True)
B = g.fcompress_B(B); print("fcompress_B %.1f s" % (time.time() - t0), B.shape, flush=True)
B, Us = g.compress_B(B); print("compress_B %.1f s" % (time.time() - t0), B.shape, flush=True)
A = g.compress_A(A, Us); print("compress_A %.1f s" % (time.time() - t0), A.shape, flush=True)
T = gtn.einsum('IJXYijklmn,XYKL->IJKLijklmn', A, B)
T = g.compress_T(T); print("compress_T %.1f s" % (time.time() - t0), T.shape, flush=True)
z4 = gtn.einsum("IJIJij,ij", T, gtn.sparse(np.full((Nphi, Nphi), 1), statistics=(0, 0)))
err = np.abs(1 - z4 / z1)
T = gtn.dense(T)
print("done: %.1f s, shape %s, stats %s, norm %.17g, z1 %r z4 %r err %.3g nnz %d" % (
    time.time() - t0, T.shape, T.statistics, T.norm, z1, z4, err, T.nnz), flush=True)
np.savez_compressed(os.path.join(HERE, "z2_initial_tensor.npz"),
                    data=np.asarray(T.data), statistics=np.array([str(s) for s in T.statistics]),
                    encoder=T.encoder, format=T.format, trace_error=err, z1=z1, z4=z4,
                    normA=normA, normB=normB, nnzB=nnzB, generator="make_z2_tensor_fast.py",
                    params=np.array([2, 1.0, 1, 1.0, 1.0, 1.0, 0.0]))
