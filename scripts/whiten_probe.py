"""gtn_chol_whiten alone: time per launch (CUDA events, 4 matrices per launch, split-K 4) and the clock64 phase stamps of
CTA 0 (factorisation / inverse) for the three kernel variants -- where the 0.7 ms of an l = 128 whitening go."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import _engine as E
from grassmanntn_b200._cabi import check, lib
from test_gpu_whiten import _gram
dev = torch.device("cuda")
for n in ([int(a) for a in sys.argv[1:]] or [48, 80, 96, 128, 160]):
    nb, ns = 4, 4
    Gs = [_gram(n, n, True, 10 + b, cond=1e4) for b in range(nb)]
    Gbuf = torch.zeros(nb * ns * n * n, dtype=torch.complex128, device=dev)
    for b, g_ in enumerate(Gs):
        parts = np.zeros((ns, n, n), dtype=complex); parts[:] = g_ / ns
        Gbuf[b * ns * n * n:(b + 1) * ns * n * n] = torch.from_numpy(parts.reshape(-1)).to(dev)
    Tbuf = torch.empty(nb * n * n, dtype=torch.complex128, device=dev)
    g_off = torch.tensor([b * ns * n * n for b in range(nb)], dtype=torch.int64, device=dev)
    t_off = torch.tensor([b * n * n for b in range(nb)], dtype=torch.int64, device=dev)
    n_d = torch.tensor([n] * nb, dtype=torch.int32, device=dev)
    kept = torch.zeros(nb, dtype=torch.int32, device=dev)
    se = int(lib.gtn_chol_whiten_scratch_elems(n))
    scratch = torch.empty(max(se * nb, 1), dtype=torch.complex128, device=dev)
    def run():
        check(lib.gtn_chol_whiten(E._ptr(Gbuf), E._ptr(Tbuf), 1, E._ptr(g_off), E._ptr(t_off), E._ptr(n_d), nb, n, ns, 1e-13,
                                  E._ptr(kept), E._ptr(scratch), None), "chol")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        run()
    e.record(); torch.cuda.synchronize()
    buf = (C.c_longlong * 8)()
    lib.gtn_debug_phase_clocks(buf)
    v = list(buf)
    print("n %3d: %.1f us per launch; cycles factor %d inverse %d; kept %s" % (n, s.elapsed_time(e) / 20 * 1e3, v[1] - v[0], v[2] - v[1], kept.cpu().tolist()))
