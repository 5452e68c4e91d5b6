"""Multi-GPU execution of the two shardable pieces of a coarse-graining step (one process per GPU,
torch.distributed; NCCL over NVLink on the B200 box, gloo for the CPU tests of the host logic).

What shards:
  * the large contraction: every output parity block is a GEMM whose row tiles are independent.
    Rank r computes rows [r*m/W, (r+1)*m/W) of every output block from the (replicated) packed
    operands; the blocks are completed with one in-place all-gather per block.  The operands of
    the big contractions are products of the isometries, so what travels is the result tile set,
    never an input tensor.
  * the sector decompositions: the E/O sector matrices of the (one or two) tensors being
    decomposed are independent problems; problem i is solved by rank i % W and its (U, s, Vh) --
    the isometries -- are broadcast from the owner.
Everything else of a step (small einsums with the singular values, norms) is replicated: it is
O(chi^2 D^2) against O(chi^4 D^2) for the sharded part.

`enable()` switches the engine to this mode; without it every rank runs the full step on its own
GPU (replicas).  The reference has no distributed mode at all.
"""
import torch
import torch.distributed as dist

_state = {"on": False, "min_flops": 2.0e9, "gemm": True, "svd": True, "fused": False}
_symm = {}          # nbytes -> (symmetric uint8 tensor, handle, ctypes array of peer pointers)
MAX_PEERS = 8       # GTN_MAX_PEERS of include/gtn_b200.h


def enable(min_flops=2.0e9, gemm=True, svd=True, fused=True):
    """shard contractions with at least `min_flops` algorithmic flops and all sector decompositions.
    fused=True (NCCL backend, <= 8 ranks, peer access): the sharded GEMM stores its result tiles straight
    into every rank's output buffer over NVLink (gtn_grouped_gemm_bcast on torch symmetric memory) instead
    of running an all-gather after the GEMM; falls back to the all-gather if symmetric memory is unavailable."""
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("grassmanntn_b200.parallel.enable(): torch.distributed is not initialised")
    _state["on"] = dist.get_world_size() > 1
    _state["min_flops"] = float(min_flops)
    _state["gemm"], _state["svd"] = bool(gemm), bool(svd)
    _state["fused"] = bool(fused and _state["on"] and dist.get_backend() == "nccl"
                           and dist.get_world_size() <= MAX_PEERS and torch.cuda.is_available())
    return _state["on"]


def fused():
    return _state["on"] and _state["fused"]


def _symm_output(nbytes, device):
    """symmetric (peer-mapped) staging buffer of the sharded contraction output, cached by size"""
    ent = _symm.get(nbytes)
    if ent is None:
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        t = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        hdl = symm_mem.rendezvous(t, group)
        w = dist.get_world_size()
        ptrs = (C.c_void_p * w)(*[int(p) for p in hdl.buffer_ptrs])
        ent = _symm[nbytes] = (t, hdl, ptrs)
    return ent


def gemm_allgather_fused(splan, A, B, out):
    """this rank's row tiles of every output block, stored by the GEMM epilogue into ALL ranks' copies of
    the output (peer memory); `out` receives the complete result.  Returns False (nothing done) when the
    symmetric buffer cannot be set up -- the caller then uses GEMM + all-gather."""
    from ._cabi import check, lib
    from ._engine import _ptr, _stream, dtype_code, prof_region
    nbytes = out.numel() * out.element_size()
    err = None
    if nbytes not in _symm:
        # first use of this size: every rank tries, then all ranks agree on the outcome (one rank falling back to
        # the NCCL all-gather while its peers wait in the symmetric-memory barrier would deadlock)
        try:
            _symm_output(nbytes, out.device)
        except Exception as exc:                  # no peer access / symmetric memory on this system
            err = exc
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=out.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            _symm.pop(nbytes, None)
            _state["fused"] = False
            import warnings
            warnings.warn("grassmanntn_b200.parallel: fused GEMM+all-gather unavailable (%s); using NCCL all-gather"
                          % (err if err is not None else "a peer could not map the buffer"))
            return False
    t, hdl, ptrs = _symm[nbytes]
    hdl.barrier(channel=0)                        # every peer has copied the previous result out of its buffer
    if splan.n and splan.tiles:
        with prof_region("grouped_gemm_bcast", 1, splan.bytes, splan.flops):
            check(lib.gtn_grouped_gemm_bcast(_ptr(A), _ptr(B), ptrs, dist.get_world_size(), dtype_code(A.dtype),
                                             _ptr(splan.dev), splan.n, splan.tiles, _stream()), "gtn_grouped_gemm_bcast")
    hdl.barrier(channel=1)                        # every rank's tiles have landed in this rank's buffer
    out.view(torch.uint8).copy_(t[:nbytes])
    return True


def disable():
    _state["on"] = False


def active():
    return _state["on"]


def world():
    return dist.get_world_size() if active() else 1


def rank():
    return dist.get_rank() if active() else 0


def row_range(m, r, w):
    """rows [lo, hi) of an m-row block owned by rank r of w; contiguous, sizes differ by at most 1"""
    base, rem = divmod(m, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def shard_groups(groups, r, w):
    """Restrict every GEMM group (dicts of _engine.GemmPlan) to the rows owned by rank r.
    Returns (sharded groups, pieces) with pieces[g] = list over ranks of (c_off, n_elems) of the
    chunk each rank writes -- the all-gather layout of block g."""
    out, pieces = [], []
    for g in groups:
        m, n = g["m"], g["n"]
        batch = g.get("batch", 1)
        if batch != 1:
            raise NotImplementedError("sharding of batched groups")
        lo, hi = row_range(m, r, w)
        gg = dict(g)
        gg["m"] = hi - lo
        gg["a_off"] = g["a_off"] + lo * g["lda"]
        gg["c_off"] = g["c_off"] + lo * g["ldc"]
        out.append(gg)
        pieces.append([(g["c_off"] + row_range(m, q, w)[0] * g["ldc"],
                        (row_range(m, q, w)[1] - row_range(m, q, w)[0]) * g["ldc"]) for q in range(w)])
    return out, pieces


def gather_blocks(buf, pieces):
    """complete every output block: rank q's chunk of block g lives at pieces[g][q] inside `buf` on
    every rank; after the call all chunks are valid everywhere.  One broadcast per (block, rank) is
    replaced by one all_gather per block when the chunks are equal-sized (the common case)."""
    w = dist.get_world_size()
    r = dist.get_rank()
    for pc in pieces:
        sizes = {n for _, n in pc}
        if len(sizes) == 1 and all(pc[q][0] == pc[0][0] + q * pc[0][1] for q in range(w)):
            n = pc[0][1]
            if n == 0:
                continue
            whole = buf[pc[0][0]: pc[0][0] + w * n]
            mine = buf[pc[r][0]: pc[r][0] + n]
            # complex tensors are gathered through their real view (NCCL has no complex dtype)
            if whole.is_complex():
                whole, mine = torch.view_as_real(whole), torch.view_as_real(mine)
            dist.all_gather_into_tensor(whole.reshape(-1), mine.reshape(-1))
        else:
            for q, (off, n) in enumerate(pc):
                if n:
                    t = buf[off: off + n]
                    dist.broadcast(torch.view_as_real(t) if t.is_complex() else t, src=q)


def owner(i):
    return i % world()


def broadcast_usv(results, nprob, device, dtype):
    """results: {i: (U, s_numpy, Vh)} for the problems this rank solved.  Returns the full list on
    every rank (isometries broadcast from their owners)."""
    import numpy as np
    w, r = dist.get_world_size(), dist.get_rank()
    # shapes of the owners' results: one small integer all-reduce (every problem has exactly one owner), no pickling
    table = torch.zeros(nprob, 5, dtype=torch.int64, device=device)
    for i, (u, s, v) in results.items():
        table[i] = torch.tensor([u.shape[0], u.shape[1], len(s), v.shape[0], v.shape[1]], dtype=torch.int64)
    dist.all_reduce(table)
    meta = {i: ((int(t[0]), int(t[1])), int(t[2]), (int(t[3]), int(t[4]))) for i, t in enumerate(table.cpu().tolist())}
    out = []
    for i in range(nprob):
        ush, ns, vsh = meta[i]
        src = i % w
        if src == r:
            U, s, Vh = results[i]
            # private contiguous copies: the solver returns views into its workspace
            U, Vh = U.contiguous().clone(), Vh.contiguous().clone()
            torch.cuda.current_stream().synchronize() if U.is_cuda else None
            st = torch.from_numpy(np.ascontiguousarray(s)).to(device)
        else:
            U = torch.empty(ush, dtype=dtype, device=device)
            Vh = torch.empty(vsh, dtype=dtype, device=device)
            st = torch.empty(ns, dtype=torch.float64, device=device)
        for t in (U, Vh):
            if t.numel():
                dist.broadcast(torch.view_as_real(t) if t.is_complex() else t, src=src)
        if ns:
            dist.broadcast(st, src=src)
        out.append((U, st.cpu().numpy(), Vh))
    return out
