"""Golden numbers of the initial-tensor compression pipeline (reference gauge2d.py:22-66, :1198-1585) for the Z2 model
of example.py's defaults, from the REAL reference's fcompress_B / compress_B / compress_A / compress_T (its einsum_ds
swapped for the oracle's vectorised restatement, exactly as in make_z2_tensor_fast.py): the A and B tensors that
get_ABtensors produces (stored sparsely: the inputs of the GPU pipeline, grassmanntn_b200.gauge2d.tensor_from_AB) and,
per stage, the shape and the norm of the compressed tensor -- gauge invariant, unlike the isometries themselves --
plus z1, z4 and the trace error.  Build container only (needs /root/reference); writes tests/golden/z2_prep.npz."""
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
cache = "/tmp/z2_AB.pkl"
if os.path.exists(cache):
    Ad, Ast, Bd, Bst = pickle.load(open(cache, "rb"))
else:
    A_, B_ = g.get_ABtensors(Nphi=2, beta=1.0, Nf=1, spacing=1.0, mass=1.0, charge=1.0, mu=0.0, Gauss=False)
    Ad, Ast, Bd, Bst = A_.data.todense(), A_.statistics, B_.data.todense(), B_.statistics
    pickle.dump((Ad, Ast, Bd, Bst), open(cache, "wb"))
Ad, Bd = np.asarray(Ad), np.asarray(Bd)
A = gtn.sparse(Ad, statistics=Ast)
B = gtn.sparse(Bd, statistics=Bst)
t0 = time.time()
rec = {}
rec["z1"] = complex(gtn.einsum("IJIJijij,jiji", B, A))
B = g.fcompress_B(B); rec["fB_shape"], rec["fB_norm"] = np.array(B.shape), float(B.norm); print("fcompress_B", B.shape, B.norm, time.time() - t0, flush=True)
B, Us = g.compress_B(B); rec["cB_shape"], rec["cB_norm"] = np.array(B.shape), float(B.norm); print("compress_B", B.shape, B.norm, flush=True)
rec["U_shapes"] = np.array([u.shape for u in Us])
A = g.compress_A(A, Us); rec["cA_shape"], rec["cA_norm"] = np.array(A.shape), float(A.norm); print("compress_A", A.shape, A.norm, flush=True)
T = gtn.einsum('IJXYijklmn,XYKL->IJKLijklmn', A, B); rec["T0_shape"], rec["T0_norm"] = np.array(T.shape), float(T.norm)
T = g.compress_T(T); rec["T_shape"], rec["T_norm"] = np.array(T.shape), float(T.norm); print("compress_T", T.shape, T.norm, flush=True)
rec["z4"] = complex(gtn.einsum("IJIJij,ij", T, gtn.sparse(np.full((2, 2), 1), statistics=(0, 0))))
rec["trace_error"] = abs(1 - rec["z4"] / rec["z1"])
print(rec, time.time() - t0, flush=True)
bc = np.argwhere(Bd != 0)
np.savez_compressed(os.path.join(HERE, "z2_prep.npz"), A=Ad, A_stats=np.array(Ast), B_shape=np.array(Bd.shape),
                    B_stats=np.array(Bst), B_coords=bc.astype(np.int16), B_vals=Bd[tuple(bc.T)], **rec)
