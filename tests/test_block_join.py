"""join_legs_block / split_legs_block (reference __init__.py:3322-3859) against golden
vectors generated from the real reference (tests/golden/make_block_join.py).  CPU: the oracle restatement,
bit-exact.  GPU: the product (one sign+permute launch per call), bit-exact."""
import itertools
import os

import numpy as np
import pytest

import gtn_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# must match tests/golden/make_block_join.py
CASES = [
    ((2, 2, 2, 2), (2, 2, 2, 2), (1, 1, -1, -1), "standard", "(ij)(kl)", (1, -1)),
    ((2, 2, 2, 2), (2, 2, 2, 2), (1, 1, -1, -1), "standard", "(ij)(kl)", (-1, 1)),
    ((2, 1, 4, 2), (2, 1, 4, 2), (1, -1, -1, 1), "matrix", "(ij)(kl)", (-1, -1)),
    ((3, 2, 2, 1), (2, 3, 1, 2), (1, 1, -1, -1), "standard", "(ij)(kl)", (1, -1)),
    ((3, 2, 2, 1), (2, 3, 1, 2), (1, -1, 1, -1), "matrix", "(ijk)(l)", (-1, 1)),
    ((2, 3, 2), (1, 2, 2), (1, 1, -1), "standard", "(ijk)", (-1,)),
    ((2, 2, 3, 2), (2, 1, 3, 2), (1, -1, 0, 0), "standard", "(ij)(kl)", (1, 0)),
    ((2, 3, 2, 2, 2), (2, 3, 1, 2, 3), (-1, 0, 1, 1, -1), "standard", "(i)(j)(klm)", (-1, 0, -1)),
    ((2, 2, 1, 2, 2), (1, 2, 2, 2, 1), (1, 1, 1, -1, -1), "matrix", "(ijk)(lm)", (1, -1)),
    ((2, 2, 2, 2, 2, 2), (2, 2, 2, 2, 2, 2), (1, 1, 1, 1, -1, -1), "standard", "(ijkl)(mn)", (-1, 1)),
]
Z = np.load(os.path.join(G, "block_join.npz"))


def _pats(stats):
    nf = sum(1 for s in stats if s in (1, -1))
    return list(itertools.product((0, 1), repeat=nf))


def _key(k, what, p):
    return "c%d_%s_%s" % (k, what, "".join(map(str, p)))


def _input_blocks(k, stats):
    return {p: Z[_key(k, "in", p)] for p in _pats(stats)}


def test_golden_matches_case_table():
    assert int(Z["n"]) == len(CASES)


@pytest.mark.parametrize("k", range(len(CASES)))
def test_oracle_join_split_block_vs_reference(k):
    ev, od, st, fmt, string, fstat = CASES[k]
    od_full = [o if s in (1, -1) else e for e, o, s in zip(ev, od, st)]
    B = O.Blocks(st, ev, od_full, _input_blocks(k, st), fmt)
    J = O.join_legs_block(B, string, fstat)
    assert J.marked_as_joined and J.format == str(Z["c%d_join_format" % k]) == fmt
    fj = [a for a, s in enumerate(fstat) if s in (1, -1)]
    assert [J.eshape[a] for a in fj] == [int(Z["c%d_join_even" % k][a]) for a in fj]
    assert [J.oshape[a] for a in fj] == [int(Z["c%d_join_odd" % k][a]) for a in fj]
    for p in _pats(fstat):
        assert np.array_equal(J.blocks[p], Z[_key(k, "join", p)]), (k, p)
    for a in fj:
        for pi in (0, 1):
            ref = Z["c%d_sgn_%d_%d" % (k, pi, a)]
            n = J.eshape[a] if pi == 0 else J.oshape[a]
            assert np.array_equal(J.sgn[(pi, a)], ref[:n]), (k, pi, a)
    Jm = J.switch_format()
    for p in _pats(fstat):
        assert np.array_equal(Jm.blocks[p], Z[_key(k, "joinsw", p)]), (k, p)
    S = O.split_legs_block(J, string, st, tuple(int(x) for x in Z["c%d_shape" % k]), ev, od)
    assert S.format == str(Z["c%d_split_format" % k]) == fmt
    for p in _pats(st):
        assert np.array_equal(S.blocks[p], Z[_key(k, "split", p)]), (k, p)
        if fmt == "standard":
            assert np.array_equal(S.blocks[p], B.blocks[p])


def test_oracle_join_block_errors():
    ev, od, st, fmt, string, fstat = CASES[0]
    B = O.Blocks(st, ev, od, _input_blocks(0, st), fmt)
    J = O.join_legs_block(B, string, fstat)
    with pytest.raises(ValueError):
        O.join_legs_block(J, "(i)(j)", fstat)                     # joined once only
    with pytest.raises(ValueError):
        O.split_legs_block(B, string, st, (4, 4, 4, 4), ev, od)   # not joined
    with pytest.raises(ValueError):
        O.join_legs_block(B, "(ij)(k)", fstat)                    # index count
    ev, od, st, fmt, string, fstat = CASES[6]
    B = O.Blocks(st, ev, [2, 1, 3, 2], _input_blocks(6, st), fmt)
    with pytest.raises(ValueError):
        O.join_legs_block(B, "(ijk)(l)", (1, 0))                  # hybrid group


# ---------------------------------------------------------------------------------------------------------------
#  product: host-built launch tables (CPU, kernel emulated by tests/test_tables_cpu.emulate) and the GPU itself
# ---------------------------------------------------------------------------------------------------------------
def _product_case(gtn, k, device):
    """the golden case as a product block tensor (all parity blocks stored)"""
    import torch
    from grassmanntn_b200 import _engine as E
    ev, od, st, fmt, string, fstat = CASES[k]
    bt = E.BT(st, ev, od, torch.complex128, fmt)
    total = 0
    for p in bt.patterns():
        bt.off[p] = total
        total += bt.block_size(p)
    buf = np.zeros(max(total, 1), dtype=np.complex128)
    for p in bt.patterns():
        buf[bt.off[p]: bt.off[p] + bt.block_size(p)] = Z[_key(k, "in", p)].ravel()
    bt.buf = torch.from_numpy(buf).to(device)
    return gtn.block._from_bt(bt, tuple(int(x) for x in Z["c%d_shape" % k]))


def _check_product_case(gtn, k, device):
    ev, od, st, fmt, string, fstat = CASES[k]
    B = _product_case(gtn, k, device)
    J = B.join_legs(string, fstat)
    assert J.marked_as_joined and J.format == fmt and J.statistics == tuple(fstat)
    fj = [a for a, s in enumerate(fstat) if s in (1, -1)]
    assert [J.even_shape[a] for a in fj] == [int(Z["c%d_join_even" % k][a]) for a in fj]
    assert [J.odd_shape[a] for a in fj] == [int(Z["c%d_join_odd" % k][a]) for a in fj]
    assert tuple(J.shape) == tuple(int(x) for x in Z["c%d_join_shape" % k])
    data = J.data
    for p in _pats(fstat):
        assert np.array_equal(data[p].cpu().numpy(), Z[_key(k, "join", p)]), (k, p)
    sgn = J.sgn
    for a in fj:
        for pi in (0, 1):
            n = J.even_shape[a] if pi == 0 else J.odd_shape[a]
            assert np.array_equal(sgn[pi][a][:n], Z["c%d_sgn_%d_%d" % (k, pi, a)][:n]), (k, pi, a)
    dsw = J.switch_format().data
    for p in _pats(fstat):
        assert np.array_equal(dsw[p].cpu().numpy(), Z[_key(k, "joinsw", p)]), (k, p)
    S = J.split_legs(string, st, tuple(int(x) for x in Z["c%d_shape" % k]), ev, od)
    assert not S.marked_as_joined and S.format == fmt and S.statistics == tuple(st)
    assert tuple(S.even_shape) == tuple(ev)
    ds = S.data
    for p in _pats(st):
        assert np.array_equal(ds[p].cpu().numpy(), Z[_key(k, "split", p)]), (k, p)
    # errors: joined once, split only joined, no contraction / decomposition of a joined tensor
    with pytest.raises(ValueError):
        J.join_legs("(i)" * J.ndim, fstat)
    with pytest.raises(ValueError):
        B.split_legs(string, st, B.shape, ev, od)
    with pytest.raises(ValueError):
        gtn.einsum("".join("abcdef"[:J.ndim]) + "->" + "".join("abcdef"[:J.ndim]), J)


@pytest.mark.parametrize("k", range(len(CASES)))
def test_product_tables_join_split_block_vs_reference(host_tables, k):
    import grassmanntn_b200 as gtn
    _check_product_case(gtn, k, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(CASES)))
def test_gpu_join_split_block_vs_reference(gtn, k):
    _check_product_case(gtn, k, "cuda")


@pytest.mark.gpu
def test_gpu_join_split_block_roundtrip_at_size(gtn):
    """size-independent property at the bench size: split(join(T)) == T bit-exactly for a standard-format
    32^4 block tensor, and the joined matrix has the norm of T"""
    g = gtn.gauge2d
    T = g.zcap(g.load_initial_tensor()).toblock()
    for _ in range(2):
        T, _ = g.trg(T, 32)
    J = T.join_legs("(ij)(kl)", (1, -1))
    assert J.even_shape == (512, 512) and abs(J.norm - T.norm) <= 1e-14 * T.norm
    S = J.split_legs("(ij)(kl)", T.statistics, T.shape, T.even_shape, T.odd_shape)
    a, b = T.data, S.data
    for p in T._bt.patterns():
        assert bool((a[p] == b[p]).all()), p
