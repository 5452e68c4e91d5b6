"""CPU: the product's host logic for the sign+permute family (dense <-> block conversion, format switch, Hermitian
conjugate, leg permutation with Grassmann signs) against the oracle, with every launch replaced by the numpy
emulation of the kernel's addressing / sign rule (conftest.host_tables).  No GPU, no product code path changed:
the emulation stands in for gtn_sign_permute only."""
import itertools

import numpy as np
import pytest
import torch

import gtn_oracle as O


def _product_block(gtn, Bo):
    """oracle Blocks -> product block tensor with host buffers (all parity blocks stored)"""
    from grassmanntn_b200 import _engine as E
    cplx = any(np.iscomplexobj(b) for b in Bo.blocks.values())
    bt = E.BT(Bo.statistics, Bo.eshape, Bo.oshape, torch.complex128 if cplx else torch.float64, Bo.format)
    total = 0
    for p in bt.patterns():
        bt.off[p] = total
        total += bt.block_size(p)
    buf = np.zeros(max(total, 1), dtype=np.complex128 if cplx else np.float64)
    for p in bt.patterns():
        buf[bt.off[p]: bt.off[p] + bt.block_size(p)] = np.asarray(Bo.blocks[p]).ravel()
    bt.buf = torch.from_numpy(buf)
    return gtn.block._from_bt(bt)


CASES = [((4, 2, 8, 4), (1, -1, -1, 1), True), ((4, 4, 2), (1, 1, -1), True), ((2, 8, 3, 4), (-1, 1, 0, -1), True),
         ((4, 4, 4, 4), (1, 1, -1, -1), False)]


@pytest.mark.parametrize("shape,stats,cplx", CASES)
def test_block_to_dense_and_format_switch(host_tables, shape, stats, cplx):
    import grassmanntn_b200 as gtn
    rng = np.random.RandomState(11)
    A = O.random_dense(shape, stats, dtype=complex if cplx else float, rng=rng, skip_trimming=True)
    Bo = O.Blocks.from_dense(A)
    B = _product_block(gtn, Bo)
    assert np.array_equal(B.todense().data.numpy(), A.data)                          # bt_to_dense (encoder gather)
    Pd = B.todense("parity-preserving").data.numpy()
    assert np.array_equal(Pd, A.force_encoder("parity-preserving").data)
    Bs, Bos = B.switch_format(), Bo.switch_format()
    data = Bs.data
    for p in itertools.product((0, 1), repeat=len(Bo.faxes)):
        assert np.array_equal(data[p].numpy(), Bos.blocks[p]), p
    assert Bs.format == "matrix" and np.array_equal(Bs.switch_format().todense().data.numpy(), A.data)


@pytest.mark.parametrize("string,shape,stats", [("ij|kl", (4, 2, 8, 4), (1, -1, -1, 1)),
                                                ("i|jk", (4, 4, 2), (1, 1, -1)),
                                                ("ijk|l", (2, 8, 3, 4), (-1, 1, 0, -1))])
def test_hconjugate_tables(host_tables, string, shape, stats):
    import grassmanntn_b200 as gtn
    rng = np.random.RandomState(5)
    A = O.random_dense(shape, stats, dtype=complex, rng=rng)            # Grassmann-even (hconjugate joins legs)
    B = _product_block(gtn, O.Blocks.from_dense(A))
    ref = O.hconjugate(A, string)
    got = B.hconjugate(string)
    assert tuple(got.statistics) == tuple(ref.statistics)
    assert np.array_equal(got.todense().data.numpy(), ref.data)


@pytest.mark.parametrize("sub,shape,stats", [("ijkl->jkli", (4, 2, 8, 4), (1, -1, -1, 1)),
                                             ("ijkl->klij", (4, 4, 4, 4), (1, 1, -1, -1)),
                                             ("ijk->kji", (4, 4, 2), (1, 1, -1)),
                                             ("ijkl->lkij", (2, 8, 3, 4), (-1, 1, 0, -1))])
def test_signed_permutation_tables(host_tables, sub, shape, stats):
    import grassmanntn_b200 as gtn
    rng = np.random.RandomState(7)
    A = O.random_dense(shape, stats, dtype=complex, rng=rng, skip_trimming=True)
    B = _product_block(gtn, O.Blocks.from_dense(A))
    ref = O.einsum(sub, A)
    got = gtn.einsum(sub, B)
    assert np.array_equal(got.todense().data.numpy(), ref.data)
