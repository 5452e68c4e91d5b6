"""Generate tests/golden/z2_initial_tensor.npz from the REAL reference (build container only).

Restates the call sequence of gauge2d.tensor_preparation (reference gauge2d.py:22-66) --
get_ABtensors (sympy Berezin integrals, ~15 min), fcompress_B, compress_B, compress_A,
T formation, compress_T -- calling the reference's own functions, for the model of
example.py's defaults: Z_2 (K=2), N_f=1, beta=m=q=a=1, mu=0.  The (A,B) stage is cached in
/tmp so a failure downstream does not repeat the sympy part.  Stores the compressed site
tensor T (shape (8,8,8,8,2,2), statistics (1,1,-1,-1,0,0)) plus the trace error and the
known-answer norms quoted in docs/_sources/schwinger.rst.txt.  The fixture is what every Z2
parity test / bench workload starts from (the initial-tensor pipeline is
out of scope and shipped as a fixture).
"""
import os, sys, time, pickle
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_harness

gtn = ref_harness.load_reference()
g = gtn.gauge2d
Nphi = 2
cache = "/tmp/z2_AB.pkl"
t0 = time.time()
if os.path.exists(cache):
    Ad, Ast, Bd, Bst = pickle.load(open(cache, "rb"))
    A = gtn.sparse(Ad, statistics=Ast)
    B = gtn.sparse(Bd, statistics=Bst)
else:
    A, B = g.get_ABtensors(Nphi=Nphi, beta=1.0, Nf=1, spacing=1.0, mass=1.0, charge=1.0, mu=0.0, Gauss=False)
    pickle.dump((A.data.todense(), A.statistics, B.data.todense(), B.statistics), open(cache, "wb"))
print("AB: %.1f s  A %s %s norm %.17g | B %s %s nnz %d norm %.17g" % (
    time.time() - t0, A.shape, A.statistics, A.norm, B.shape, B.statistics, B.nnz, B.norm), flush=True)
normA, normB, nnzB = float(A.norm), float(B.norm), int(B.nnz)
z1 = gtn.einsum("IJIJijij,jiji", B, A)
B = g.fcompress_B(B)
B, Us = g.compress_B(B)
A = g.compress_A(A, Us)
T = gtn.einsum('IJXYijklmn,XYKL->IJKLijklmn', A, B)
T = g.compress_T(T)
z4 = gtn.einsum("IJIJij,ij", T, gtn.sparse(np.full((Nphi, Nphi), 1), statistics=(0, 0)))
err = np.abs(1 - z4 / z1)
T = gtn.dense(T)
print("done: %.1f s, shape %s, stats %s, norm %.17g, z1 %r z4 %r err %.3g" % (
    time.time() - t0, T.shape, T.statistics, T.norm, z1, z4, err), flush=True)
np.savez_compressed(os.path.join(HERE, "z2_initial_tensor.npz"),
                    data=np.asarray(T.data), statistics=np.array([str(s) for s in T.statistics]),
                    encoder=T.encoder, format=T.format, trace_error=err, z1=z1, z4=z4,
                    normA=normA, normB=normB, nnzB=nnzB,
                    params=np.array([2, 1.0, 1, 1.0, 1.0, 1.0, 0.0]))
