"""Golden numbers for the long / large chains of tests/test_chains.py (run in the build container):

  stepgraph_chain   12 TRG steps (gauge2d_block.trg, dcut 16) of the REAL reference on a perturbed Z2 site tensor
                    (zcap of the Z2 fixture + 5 % of a seeded random Grassmann-even tensor: no exact multiplets, so no
                    truncation tie-breaking) -- long enough for the product to reach its recorded whole-step CUDA
                    graph, whose results are then compared with the reference step by step.
  oracle_trg_chi64  4 TRG steps at chi = 64 on the Z2 tensor with the ORACLE PORT (oracle/gtn_oracle.py trg_block; the
                    real reference's element-wise Python loops need hours at D = 64).  The port itself is pinned
                    against the real reference at chi <= 32 by tests/test_z2_golden.py.

  atrg_chain        8 ATRG steps (alternating atrg2dx / atrg2dy like example.py:178-188, Dcut 16) of the REAL reference on the
                    same perturbed tensor: golden for the sharded ATRG chain (tests/test_sharded_gloo.py, test_gpu_sharded.py)

  python tests/golden/make_chain_goldens.py [stepgraph] [chi64] [atrg]
"""
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import gtn_oracle as O  # noqa: E402
from threadpoolctl import threadpool_limits  # noqa: E402

OUT = os.path.join(HERE, "chains.npz")
out = dict(np.load(OUT)) if os.path.exists(OUT) else {}
want = sys.argv[1:] or ["stepgraph", "chi64", "atrg"]
z = np.load(os.path.join(HERE, "z2_initial_tensor.npz"))
stats6 = tuple(int(s) for s in z["statistics"])
BC = "anti-periodic"


def perturbed_tensor():
    T0 = O.zcap(O.Dense(z["data"], stats6))
    rng = np.random.RandomState(20261017)
    R = O.random_dense(T0.shape, T0.statistics, dtype=complex, rng=rng)
    R.data = R.data - 0.5 * (1 + 1j) * (R.data != 0)            # centred
    data = T0.data + 0.05 * np.linalg.norm(T0.data) / np.linalg.norm(R.data) * R.data
    return data, T0.statistics


if "stepgraph" in want:
    import ref_harness
    gtn = ref_harness.load_reference()
    data, st = perturbed_tensor()
    T = gtn.dense(data, statistics=st).toblock()
    mod = gtn.gauge2d_block
    rec, logNorm = [], 0.0
    t0 = time.time()
    with threadpool_limits(limits=1):
        for i in range(12):
            T, Tn = mod.trg(T, 16, iternum=i)[:2]
            logNorm = 2 * logNorm + math.log(Tn)
            F = (mod.logZ(T.copy(), BC) + logNorm) / 2 ** (i + 1)
            rec.append([Tn, F.real, F.imag, T.effective_shape[0], T.effective_shape[1]])
            print("stepgraph", i, Tn, F, T.effective_shape, "%.1f s" % (time.time() - t0), flush=True)
    out["stepgraph_input"] = data
    out["stepgraph_stats"] = np.asarray(st)
    out["stepgraph_chain"] = np.array(rec, dtype=float)
    np.savez_compressed(OUT, **out)

if "atrg" in want:
    import ref_harness
    gtn = ref_harness.load_reference()
    data, st = perturbed_tensor()
    T = gtn.dense(data, statistics=st).toblock()
    mod = gtn.gauge2d_block
    rec, logNorm = [], 0.0
    cgxfirst = T.shape[0] > T.shape[1]
    t0 = time.time()
    with threadpool_limits(limits=1):
        for i in range(8):
            use_x = (i % 2 == 0) == cgxfirst
            fn = mod.atrg2dx if use_x else mod.atrg2dy
            T, Tn = fn(T, T, 16, iternum=i)[:2]
            logNorm = 2 * logNorm + math.log(Tn)
            F = (mod.logZ(T.copy(), BC) + logNorm) / 2 ** (i + 1)
            rec.append([Tn, F.real, F.imag, T.effective_shape[0], T.effective_shape[1], 1.0 if use_x else 0.0])
            print("atrg", i, "x" if use_x else "y", Tn, F, T.effective_shape, "%.1f s" % (time.time() - t0), flush=True)
    out["atrg_chain"] = np.array(rec, dtype=float)
    np.savez_compressed(OUT, **out)

if "chi64" in want:
    B = O.Blocks.from_dense(O.zcap(O.Dense(z["data"], stats6)))
    rec, logNorm = [], 0.0
    t0 = time.time()
    for i in range(4):
        B, Tn = O.trg_block(B, 64)
        logNorm = 2 * logNorm + math.log(Tn)
        F = (O.logZ(B.todense(), BC, block_format=True) + logNorm) / 2 ** (i + 1)
        rec.append([Tn, F.real, F.imag, B.effective_shape[0], B.effective_shape[1]])
        print("chi64", i, Tn, F, B.effective_shape, "%.1f s" % (time.time() - t0), flush=True)
    out["oracle_trg_chi64"] = np.array(rec, dtype=float)
    np.savez_compressed(OUT, **out)
