"""On-disk format for long coarse-graining runs: tensor checkpoints and a per-step observable log.

The reference keeps nothing on disk (no checkpoint / resume; example.py prints one line
per step and loses the tensor when the process ends).  SURVEY.md section 8(f) row 4 asks for a format that makes
chi = 128..256 runs resumable and comparable:

* `save_tensor` / `load_tensor` -- one `.npz` per tensor holding the parity-blocked device buffer exactly as it
  lives in HBM (`_engine.BT`: one contiguous buffer + block offsets; only the stored, i.e. even-parity, blocks),
  its statistics / even / odd dimensions / format, the container kind (dense | block, encoder) and free-form
  metadata.  Loading restores the same tensor bits; the continuation agrees with an uninterrupted run to rounding:
  the kernels themselves are bitwise reproducible (tests/test_gpu_edges.py::test_chain_is_bitwise_reproducible), but
  the adaptive state of the truncated SVD -- iteration counts, recorded graphs -- is not checkpointed, so a resumed
  run may iterate a different number of times than the uninterrupted one.
* `RunLog` -- JSON-lines file, one record per coarse-graining step with the columns example.py prints
  (reference example.py:158-196): process, volume, F (real, imag), shape, trace error, Tnorm, seconds.
* `gauge2d.coarse_grain(..., log=, checkpoint_dir=, resume=)` uses both.

Host-side plumbing only (numpy + json); nothing here launches a kernel.
"""
import json
import os
import time

import numpy as np
import torch

from . import _engine

FORMAT_VERSION = 1


def _bt_of(T):
    from . import block, dense
    if isinstance(T, block):
        return T._bt, "block", None
    if isinstance(T, dense):
        return T._get_bt(), "dense", T.encoder
    raise TypeError("save_tensor: expected a grassmanntn_b200 dense or block tensor")


def pack_bt(bt):
    """BT -> dict of numpy arrays (the .npz members)."""
    pats = list(bt.off)
    nf = len(bt.faxes)
    return dict(
        version=np.int64(FORMAT_VERSION),
        stats=np.asarray(bt.stats, dtype=np.int64),
        e=np.asarray(bt.e, dtype=np.int64),
        o=np.asarray(bt.o, dtype=np.int64),
        fmt=np.str_(bt.fmt),
        dtype=np.str_("complex128" if bt.dtype == torch.complex128 else "float64"),
        patterns=np.asarray(pats, dtype=np.int8).reshape(len(pats), nf),
        offsets=np.asarray([bt.off[p] for p in pats], dtype=np.int64),
        zero=np.asarray(sorted(bt.zero), dtype=np.int8).reshape(len(bt.zero), nf),
        buf=bt.buf.detach().cpu().numpy(),
    )


def unpack_bt(z, device=None):
    """inverse of pack_bt; the buffer goes to `device` (default: the current CUDA device)."""
    if int(z["version"]) != FORMAT_VERSION:
        raise ValueError("checkpoint format version %d is not supported" % int(z["version"]))
    dt = torch.complex128 if str(z["dtype"]) == "complex128" else torch.float64
    bt = _engine.BT(tuple(int(s) for s in z["stats"]), [int(x) for x in z["e"]], [int(x) for x in z["o"]], dt, str(z["fmt"]))
    for p, off in zip(z["patterns"], z["offsets"]):
        bt.off[tuple(int(x) for x in p)] = int(off)
    bt.zero = {tuple(int(x) for x in p) for p in z["zero"]}
    dev = device if device is not None else _engine.require_cuda()
    bt.buf = torch.from_numpy(np.ascontiguousarray(z["buf"])).to(dev)
    if bt.buf.dtype != dt:
        raise ValueError("checkpoint buffer dtype %s does not match its header %s" % (bt.buf.dtype, dt))
    need = max([o_ + bt.block_size(p) for p, o_ in bt.off.items()] + [0])
    if bt.buf.numel() < need:
        raise ValueError("checkpoint buffer is shorter (%d) than its block table needs (%d)" % (bt.buf.numel(), need))
    return bt


class HostTensor:
    """A tensor parked in (pinned) host memory in the storage layout of the device: the parity-block buffer plus its
    block table.  `to_host` / `from_host` move it with ONE asynchronous copy each way -- the host-side format of a
    long run that streams site tensors between host and device (and of bench.py's end-to-end measurement)."""
    __slots__ = ("buf", "stats", "e", "o", "fmt", "dtype", "off", "zero", "kind", "encoder", "shape")

    @property
    def nbytes(self):
        return self.buf.numel() * self.buf.element_size()


def to_host(T, out=None, pinned=True):
    """device tensor -> HostTensor (reusing `out`'s pinned buffer when its size fits).  The copy is enqueued on the
    current stream; synchronise before reading `out.buf`."""
    bt, kind, encoder = _bt_of(T)
    h = out if out is not None else HostTensor()
    if out is None or out.buf.numel() != bt.buf.numel() or out.buf.dtype != bt.buf.dtype:
        h.buf = torch.empty(bt.buf.numel(), dtype=bt.buf.dtype)
        if pinned and torch.cuda.is_available():
            h.buf = h.buf.pin_memory()
    h.buf.copy_(bt.buf, non_blocking=True)
    h.stats, h.e, h.o, h.fmt, h.dtype = tuple(bt.stats), tuple(bt.e), tuple(bt.o), bt.fmt, bt.dtype
    h.off, h.zero = dict(bt.off), set(bt.zero)
    h.kind, h.encoder, h.shape = kind, encoder, tuple(T.shape)
    return h


def from_host(h, device=None, out_buf=None):
    """HostTensor -> device tensor of the container kind it was taken from (one asynchronous H2D copy; into the device
    buffer `out_buf` when given -- a streaming caller rotates its own buffers instead of allocating per step)."""
    from . import block, dense
    bt = _engine.BT(h.stats, h.e, h.o, h.dtype, h.fmt)
    bt.off, bt.zero = dict(h.off), set(h.zero)
    dev = device if device is not None else _engine.require_cuda()
    if out_buf is not None and out_buf.numel() == h.buf.numel() and out_buf.dtype == h.buf.dtype:
        out_buf.copy_(h.buf, non_blocking=True)
        bt.buf = out_buf
    else:
        bt.buf = h.buf.to(dev, non_blocking=True)
    if h.kind == "block":
        return block._from_bt(bt, h.shape)
    return dense._from_bt(bt, h.encoder or "canonical")


_stream_outs = {}


def stream_steps(host_inputs, step, prepare=None, n_out=2):
    """Run `step(T) -> (T', value)` over a sequence of HostTensors, STREAMING: the H2D copy of input i + 1 and the D2H
    copy of result i - 1 overlap the computation of step i (one copy-in stream, the caller's current stream for the
    computation, one copy-out stream; inputs prefetched one step ahead, `n_out` pinned result buffers in rotation).
    Every step still moves its own input from pinned host memory and its own result back -- only the waiting is gone
    (PCIe is full duplex: 2 x 2 GiB per chi = 128 step hide behind the 0.36 s of kernels).
    prepare(T) (optional) is applied to every uploaded tensor before `step` (e.g. restoring shard metadata).
    Returns [(HostTensor of T', value)] in order; the HostTensors of steps older than n_out are reused (also by the
    next call: copy out what has to outlive it)."""
    cur = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    # pinned result buffers are kept between calls (page-locking 2 GiB takes longer than the step it would hide)
    outs = _stream_outs.setdefault(n_out, [None] * n_out)
    out_done = [None] * n_out
    results = []
    n = len(host_inputs)
    dev_in = [None, None]                    # two device input buffers in rotation (no allocation per step)
    in_free = [None, None]                   # event: the step that consumed the buffer has finished

    def upload(i):
        k = i % 2
        h = host_inputs[i]
        with torch.cuda.stream(s_in):
            if dev_in[k] is None or dev_in[k].numel() != h.buf.numel() or dev_in[k].dtype != h.buf.dtype:
                dev_in[k] = torch.empty(h.buf.numel(), dtype=h.buf.dtype, device=_engine.require_cuda())
                dev_in[k].record_stream(cur)
            if in_free[k] is not None:
                s_in.wait_event(in_free[k])
            X = from_host(h, out_buf=dev_in[k])
            ev = torch.cuda.Event()
            ev.record(s_in)
        return X, ev
    nxt = upload(0) if n else None
    for i in range(n):
        X, ev = nxt
        nxt = upload(i + 1) if i + 1 < n else None          # prefetch: overlaps this step's kernels
        cur.wait_event(ev)
        if prepare is not None:
            prepare(X)
        Y, val = step(X)
        _bt_of(Y)[0].buf                                     # (a result that is an unwritten permutation is written here,
        done = torch.cuda.Event()                            #  on the compute stream, while its source is still valid)
        done.record(cur)
        in_free[i % 2] = done
        k = i % n_out
        if out_done[k] is not None:
            out_done[k].synchronize()                        # the pinned buffer's previous copy-out has landed
        s_out.wait_event(done)
        with torch.cuda.stream(s_out):
            outs[k] = to_host(Y, out=outs[k])
            _bt_of(Y)[0].buf.record_stream(s_out)
            out_done[k] = torch.cuda.Event()
            out_done[k].record(s_out)
        results.append((outs[k], val))
    s_out.synchronize()
    return results


def save_tensor(path, T, **meta):
    """Write T (dense or block) and JSON-serialisable metadata to `path` (.npz), atomically."""
    bt, kind, encoder = _bt_of(T)
    arrays = pack_bt(bt)
    arrays["kind"] = np.str_(kind)
    arrays["encoder"] = np.str_(encoder or "")
    arrays["shape"] = np.asarray(T.shape, dtype=np.int64)
    arrays["meta"] = np.str_(json.dumps(meta))
    tmp = path + ".tmp.npz"
    with open(tmp, "wb") as f:
        np.savez(f, **arrays)
    os.replace(tmp, path)
    return path


def load_tensor(path, device=None):
    """-> (T, meta): the tensor as the container kind it was saved from, bit-identical."""
    from . import block, dense
    with np.load(path, allow_pickle=False) as z:
        bt = unpack_bt(z, device)
        kind, encoder = str(z["kind"]), str(z["encoder"])
        shape = tuple(int(x) for x in z["shape"])
        meta = json.loads(str(z["meta"]))
    if kind == "block":
        return block._from_bt(bt, shape), meta
    return dense._from_bt(bt, encoder or "canonical"), meta


class RunLog:
    """Append-only JSON-lines log of a coarse-graining run."""

    def __init__(self, path, truncate=False):
        self.path = path
        if truncate and os.path.exists(path):
            os.remove(path)

    def write(self, **rec):
        rec.setdefault("time", time.time())
        for k, v in list(rec.items()):
            if isinstance(v, complex):
                rec[k] = [v.real, v.imag]
            elif isinstance(v, (np.floating, np.integer)):
                rec[k] = v.item()
            elif isinstance(v, tuple):
                rec[k] = list(v)
        with open(self.path, "a") as f:
            f.write(json.dumps(rec) + "\n")

    def read(self):
        if not os.path.exists(self.path):
            return []
        with open(self.path) as f:
            return [json.loads(line) for line in f if line.strip()]

    def truncate_to(self, nrecords):
        """keep the first `nrecords` records (a resumed run continues from a checkpoint that may be older than the
        last logged step: the steps after it are run, and logged, again)"""
        recs = self.read()[:nrecords]
        with open(self.path, "w") as f:
            for r in recs:
                f.write(json.dumps(r) + "\n")


def step_path(directory, step):
    return os.path.join(directory, "step_%04d.npz" % step)


def latest_step(directory):
    """-> (step, path) of the newest complete checkpoint in `directory`, or (None, None)."""
    best = (None, None)
    if not os.path.isdir(directory):
        return best
    for name in os.listdir(directory):
        if name.startswith("step_") and name.endswith(".npz") and ".tmp" not in name:
            try:
                s = int(name[5:-4])
            except ValueError:
                continue
            if best[0] is None or s > best[0]:
                best = (s, os.path.join(directory, name))
    return best
