"""Small driver for ncu captures (scripts, not product): runs the three kernel families once each
at sizes large enough to leave L2 -- sign+permute 'ijkl->jkli' on a dense D=96? no: D=64 (256 MiB)
and D=128 (4 GiB) complex128 tensors, one 4096^3 ZGEMM on the DMMA kernel, one TRG step at chi=32."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np, torch
import grassmanntn_b200 as gtn
from grassmanntn_b200 import _engine as E, _ops
which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda")
if which in ("all", "permute"):
    D = int(os.environ.get("GTN_D", "128"))
    x = torch.rand(D ** 4, dtype=torch.float64, device=dev).to(torch.complex128).view(D, D, D, D)
    bt = gtn.dense(x, statistics=(1, 1, -1, -1))._get_bt()
    for _ in range(3):
        r = _ops.einsum_bt('ijkl->jkli', [bt]); r.buf        # (.buf: a pure permutation is written on demand)
    torch.cuda.synchronize()
if which in ("all", "gemm"):
    N = 4096
    a = torch.randn(N, N, dtype=torch.complex128, device=dev)
    b = torch.randn(N, N, dtype=torch.complex128, device=dev)
    for _ in range(3):
        c = E.gemm(a.view(-1), b.view(-1), N, N, N)
    torch.cuda.synchronize()
if which in ("all", "trg"):
    g = gtn.gauge2d
    X = g.zcap(g.load_initial_tensor()).toblock()      # the bench workload: Z2 site tensor, chi = 32
    for _ in range(5):
        X, n = gtn.gauge2d.trg(X, 32)
    torch.cuda.synchronize()
    print("Tnorm", n)
