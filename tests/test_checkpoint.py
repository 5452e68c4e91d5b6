"""Checkpoint / run-log format.  CPU: the serialisation of the block storage
and the JSON-lines log (host plumbing, no kernel).  GPU: a resumed coarse-graining run continues bit-identically."""
import json
import os

import numpy as np
import pytest
import torch


def _fake_bt():
    """a BT with a host buffer: stats (1,-1,0), even dims (3,2,2), odd dims (2,1,-), even blocks stored"""
    from grassmanntn_b200 import _engine as E
    bt = E.BT((1, -1, 0), (3, 2, 2), (2, 1, 2), torch.complex128, "matrix")
    total = 0
    for p in [(0, 0), (1, 1), (0, 1)]:
        bt.off[p] = total
        total += bt.block_size(p)
    bt.zero = {(0, 1)}
    rng = np.random.RandomState(1)
    bt.buf = torch.from_numpy(rng.rand(total) + 1j * rng.rand(total))
    return bt


def test_pack_unpack_roundtrip_bit_exact(tmp_path):
    from grassmanntn_b200 import checkpoint as ck
    bt = _fake_bt()
    arrays = ck.pack_bt(bt)
    path = tmp_path / "t.npz"
    np.savez(path, **arrays)
    with np.load(path, allow_pickle=False) as z:
        bt2 = ck.unpack_bt(z, device="cpu")
    assert bt2.stats == bt.stats and bt2.e == bt.e and bt2.o == bt.o and bt2.fmt == "matrix"
    assert bt2.off == bt.off and bt2.zero == bt.zero and bt2.dtype == bt.dtype
    assert torch.equal(bt2.buf, bt.buf)
    assert bt2.key() == bt.key()


def test_unpack_rejects_short_buffer_and_bad_version(tmp_path):
    from grassmanntn_b200 import checkpoint as ck
    arrays = ck.pack_bt(_fake_bt())
    short = dict(arrays, buf=arrays["buf"][:-1])
    with pytest.raises(ValueError):
        ck.unpack_bt(short, device="cpu")
    with pytest.raises(ValueError):
        ck.unpack_bt(dict(arrays, version=np.int64(99)), device="cpu")


def test_runlog_and_latest_step(tmp_path):
    from grassmanntn_b200 import checkpoint as ck
    log = ck.RunLog(str(tmp_path / "run.jsonl"))
    log.write(process="_ini", vol=1, F=complex(1.5, -0.25), shape=(8, 8), Tnorm=None, err=None)
    log.write(process="_trg", vol=2, F=complex(1.25, 0.0), shape=(32, 32), Tnorm=np.float64(3.0), err=1e-3)
    recs = log.read()
    assert [r["process"] for r in recs] == ["_ini", "_trg"]
    assert recs[0]["F"] == [1.5, -0.25] and recs[1]["shape"] == [32, 32] and recs[1]["Tnorm"] == 3.0
    assert all(json.dumps(r) for r in recs)
    d = tmp_path / "ck"
    assert ck.latest_step(str(d)) == (None, None)
    os.makedirs(d)
    for s in (1, 2, 10):
        open(ck.step_path(str(d), s), "wb").close()
    open(os.path.join(d, "step_0011.npz.tmp.npz"), "wb").close()
    assert ck.latest_step(str(d)) == (10, ck.step_path(str(d), 10))
    log2 = ck.RunLog(str(tmp_path / "run.jsonl"), truncate=True)
    assert log2.read() == []


@pytest.mark.gpu
@pytest.mark.parametrize("method,cut", [("trg", 32), ("atrg", 16)])
def test_gpu_checkpoint_and_resume(gtn, tmp_path, method, cut):
    """2 steps with checkpoints, then resume for the 3rd: the reloaded tensor has the saved bits, the resumed run
    reproduces the uninterrupted 3-step run (1e-9: the resumed run starts with fresh iteration memory, so its
    truncated SVDs may iterate a different number of times) and refuses other run parameters."""
    from grassmanntn_b200 import checkpoint as ck
    g = gtn.gauge2d
    T0 = g.zcap(g.load_initial_tensor()).toblock()
    Ta, ra = g.coarse_grain(T0, cgsteps=3, dcut=cut, method=method, error_test=True)
    d, logp = str(tmp_path / "ck"), str(tmp_path / "log.jsonl")
    T2, r2 = g.coarse_grain(T0, cgsteps=2, dcut=cut, method=method, error_test=True, checkpoint_dir=d, log=logp)
    assert ck.latest_step(d)[0] == 2
    Tl, meta = ck.load_tensor(ck.step_path(d, 2))
    assert type(Tl) is type(T2) and Tl.shape == T2.shape and Tl.statistics == T2.statistics
    assert Tl._bt.key() == T2._bt.key() and bool((Tl._bt.buf == T2._bt.buf).all())          # bit-exact
    assert meta["dcut"] == cut and len(meta["records"]) == 3
    Tb, rb = g.coarse_grain(T0, cgsteps=3, dcut=cut, method=method, error_test=True, checkpoint_dir=d, log=logp,
                            resume=True)
    assert len(rb) == len(ra) == 4
    for x, y in zip(r2, rb[:3]):
        assert x["F"] == y["F"] and x["Tnorm"] == y["Tnorm"]                                # carried, not recomputed
    for x, y in zip(ra, rb):
        assert abs(x["F"] - y["F"]) <= 1e-9 * abs(x["F"])
        if x["Tnorm"] is not None:
            assert abs(x["Tnorm"] - y["Tnorm"]) <= 1e-9 * abs(x["Tnorm"])
    assert Ta._bt.key() == Tb._bt.key()
    assert abs(Ta.norm - Tb.norm) <= 1e-9
    recs = ck.RunLog(logp).read()
    assert [r["process"] for r in recs][0] == "_ini" and len(recs) == 4
    with pytest.raises(ValueError):
        g.coarse_grain(T0, cgsteps=3, dcut=8, method=method, checkpoint_dir=d, resume=True)


@pytest.mark.gpu
def test_gpu_stream_steps_equals_serial(gtn):
    """checkpoint.stream_steps (H2D of step i + 1 and D2H of step i - 1 overlapping the kernels of step i, three
    streams, rotating pinned buffers) returns what the same steps give one after the other"""
    import torch
    from grassmanntn_b200 import checkpoint as ck
    g = gtn.gauge2d
    T = g.zcap(g.load_initial_tensor()).toblock()
    chain = []
    for _ in range(4):                                   # four different inputs: the first tensors of the chi = 16 chain
        chain.append(ck.to_host(T))
        T, _ = g.trg(T, 16)
    torch.cuda.synchronize()
    serial = []
    for h in chain:
        Y, tn = g.trg(ck.from_host(h), 16)
        serial.append((float(tn), float(Y.norm), Y._bt.key()))
    res = ck.stream_steps(chain, lambda X: g.trg(X, 16), n_out=4)
    assert len(res) == 4
    for (h, tn), (tn_s, nrm_s, key_s) in zip(res, serial):
        assert abs(float(tn) - tn_s) <= 1e-12 * tn_s
        Y = ck.from_host(h)
        assert Y._bt.key() == key_s and abs(float(Y.norm) - nrm_s) <= 1e-12 * nrm_s
    # rotation with fewer pinned buffers than steps: the last results are intact
    res2 = ck.stream_steps(chain, lambda X: g.trg(X, 16), n_out=2)
    assert abs(float(res2[-1][1]) - serial[-1][0]) <= 1e-12 * serial[-1][0]
    assert abs(float(ck.from_host(res2[-1][0]).norm) - serial[-1][1]) <= 1e-12 * serial[-1][1]
