"""Golden per-step observables of the coarse-graining loops on the Z2 initial tensor, from the REAL
reference (gauge2d / gauge2d_block through oracle/ref_harness.py).  Mirrors example.py:156-196:
zcap, then `cgsteps` steps of trg / atrg, recording (Tnorm, trace error, Re F, Im F, shape).
Block configs use the fixed power_block (see ref_harness).  Slow (Python loops in the reference):
run in the background;  python tests/golden/make_z2_cg.py [config ...]"""
import os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_harness
from threadpoolctl import threadpool_limits

gtn = ref_harness.load_reference()
z = np.load(os.path.join(HERE, "z2_initial_tensor.npz"))
T0 = gtn.dense(z["data"], statistics=tuple(int(s) for s in z["statistics"]))
CONFIGS = {
    "block_trg_chi32": ("block", "trg", 32, 3),
    "block_trg_chi25": ("block", "trg", 25, 2),
    "block_atrg_chi16": ("block", "atrg", 16, 3),
    "block_atrg_chi32": ("block", "atrg", 32, 4),      # BASELINE config 1: example.py's default (ATRG, --Dcutxy 32)
    "dense_trg_chi16": ("dense", "trg", 16, 2),
    "dense_atrg_chi8": ("dense", "atrg", 8, 3),
}
want = [w for w in sys.argv[1:] if w in CONFIGS] or ([] if sys.argv[1:] else list(CONFIGS))
out_path = os.path.join(HERE, "z2_cg.npz")
out = dict(np.load(out_path)) if os.path.exists(out_path) else {}
bc = "anti-periodic"
with threadpool_limits(limits=1):
    for name in want:
        fmt, algo, cut, steps = CONFIGS[name]
        mod = gtn.gauge2d_block if fmt == "block" else gtn.gauge2d
        T = mod.zcap(T0.copy() if fmt == "dense" else T0.toblock())
        logNorm = 0.0
        F = mod.logZ(T.copy(), bc) + logNorm
        rec = [[0.0, 0.0, F.real, F.imag, T.shape[0], T.shape[1]]]
        cgxfirst = T.shape[0] > T.shape[1]
        t0 = time.time()
        for i in range(steps):
            if algo == "trg":
                T, Tn, err = mod.trg(T, cut, iternum=i, error_test=True)
            else:
                use_x = (i % 2 == 0) == cgxfirst
                fn = mod.atrg2dx if use_x else mod.atrg2dy
                T, Tn, err = fn(T, T, cut, iternum=i, error_test=True)
            logNorm = 2 * logNorm + np.log(Tn)
            F = (mod.logZ(T.copy(), bc) + logNorm) / 2 ** (i + 1)
            shp = T.effective_shape if fmt == "block" else T.shape
            rec.append([Tn, err, F.real, F.imag, shp[0], shp[1]])
            print(name, i, Tn, err, F, shp, "%.1f s" % (time.time() - t0), flush=True)
        out[name] = np.array(rec, dtype=float)
        np.savez_compressed(out_path, **out)

# ---- flavour coarse-graining (example.py:144-154 with --Nf 2): hotrg3dz(T,T,Zcut) -> zcap -> trg
if not sys.argv[1:] or "hotrg" in sys.argv[1:]:
    with threadpool_limits(limits=1):
        for cut in (8, 16):
            T, Tn, err = gtn.gauge2d.hotrg3dz(T0.copy(), T0.copy(), cut, iternum=0, error_test=True)
            logNorm = np.log(Tn)
            Tc = gtn.gauge2d.zcap(T)
            F = gtn.gauge2d.logZ(Tc.copy(), bc) + logNorm
            rec = [[Tn, err, F.real, F.imag, T.shape[0], T.shape[1]]]
            for i in range(2):
                Tc, Tn2, err2 = gtn.gauge2d.trg(Tc, cut, iternum=i, error_test=True)
                logNorm = 2 * logNorm + np.log(Tn2)
                F = (gtn.gauge2d.logZ(Tc.copy(), bc) + logNorm) / 2 ** (i + 1)
                rec.append([Tn2, err2, F.real, F.imag, Tc.shape[0], Tc.shape[1]])
            out["dense_hotrg_chi%d" % cut] = np.array(rec, dtype=float)
            print("hotrg", cut, rec, flush=True)
            np.savez_compressed(out_path, **out)
