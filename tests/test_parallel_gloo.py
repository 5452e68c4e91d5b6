"""CPU, world_size 2, gloo: the host logic of the multi-GPU path (grassmanntn_b200/parallel.py) --
row partition of the output blocks, the per-block in-place all-gather and the broadcast of the
isometries.  The GEMM itself is emulated with numpy here (the CUDA kernels need a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grassmanntn_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.enable(min_flops=0.0)
        assert not parallel.fused()          # peer-store fusion needs NCCL + CUDA; gloo uses GEMM + all-gather
        rng = np.random.RandomState(0)                      # same data on every rank
        # two output blocks (m x n) = A (m x k) B (k x n) inside shared buffers, odd sizes on purpose
        shapes = [(7, 5, 6), (4, 3, 8)]
        groups, a_off, b_off, c_off = [], 0, 0, 0
        for m, n, k in shapes:
            groups.append(dict(a_off=a_off, b_off=b_off, c_off=c_off, lda=k, ldb=n, ldc=n, m=m, n=n, k=k))
            a_off += m * k
            b_off += k * n
            c_off += m * n
        A = rng.rand(a_off) + 1j * rng.rand(a_off)
        B = rng.rand(b_off) + 1j * rng.rand(b_off)
        full = np.zeros(c_off, dtype=complex)
        for g in groups:
            full[g["c_off"]: g["c_off"] + g["m"] * g["n"]] = (
                A[g["a_off"]: g["a_off"] + g["m"] * g["k"]].reshape(g["m"], g["k"])
                @ B[g["b_off"]: g["b_off"] + g["k"] * g["n"]].reshape(g["k"], g["n"])).ravel()
        sg, pieces = parallel.shard_groups(groups, rank, world)
        buf = torch.zeros(c_off, dtype=torch.complex128)
        for g in sg:                                        # this rank's rows only
            if g["m"] == 0:
                continue
            c = (A[g["a_off"]: g["a_off"] + g["m"] * g["k"]].reshape(g["m"], g["k"])
                 @ B[g["b_off"]: g["b_off"] + g["k"] * g["n"]].reshape(g["k"], g["n"]))
            buf[g["c_off"]: g["c_off"] + g["m"] * g["n"]] = torch.from_numpy(c.ravel())
        parallel.gather_blocks(buf, pieces)
        ok1 = np.allclose(buf.numpy(), full, rtol=0, atol=1e-13)
        # isometry broadcast: problem i owned by rank i % world
        res = {}
        for i in range(3):
            if parallel.owner(i) == rank:
                r2 = np.random.RandomState(10 + i)
                res[i] = (torch.from_numpy(r2.rand(4 + i, 2) + 0j), r2.rand(2), torch.from_numpy(r2.rand(2, 5) + 0j))
        out = parallel.broadcast_usv(res, 3, torch.device("cpu"), torch.complex128)
        ok2 = True
        for i in range(3):
            r2 = np.random.RandomState(10 + i)
            u, s, v = r2.rand(4 + i, 2) + 0j, r2.rand(2), r2.rand(2, 5) + 0j
            ok2 &= np.array_equal(out[i][0].numpy(), u) and np.array_equal(out[i][1], s) and np.array_equal(out[i][2].numpy(), v)
        q.put((rank, bool(ok1), bool(ok2)))
    finally:
        dist.destroy_process_group()


def test_row_range_partition():
    for m in (0, 1, 7, 64, 100):
        for w in (1, 2, 3, 8):
            rows = [parallel.row_range(m, r, w) for r in range(w)]
            assert rows[0][0] == 0 and rows[-1][1] == m
            assert all(rows[i][1] == rows[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in rows]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_sharded_contraction_and_isometry_broadcast_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(2)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res
