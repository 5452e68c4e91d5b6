"""CPU: the launch tables the host builds for the sign+permute kernel, executed by a numpy
emulation of the kernel's addressing/sign rule (csrc/gtn_permute.cu), must reproduce the oracle."""
import numpy as np

import gtn_oracle as O
from grassmanntn_b200 import _engine as E


def emulate(fields, tabs, src, dst, scale=1.0):
    sizes = [len(t) for t in tabs]
    for idx in np.ndindex(*sizes):
        i = fields["in_base"]
        o = fields["out_base"]
        e = fields["const"]
        Macc = 0
        for x, ix in enumerate(idx):
            en = tabs[x][ix]
            i += int(en["in_off"])
            o += int(en["out_off"])
            P = int(en["P"]) & 0x0FFFFFFF
            e ^= (int(en["P"]) >> 31) ^ (bin(P & Macc).count("1") & 1)
            Macc ^= int(en["M"])
        v = src[i]
        if fields["conj"]:
            v = np.conj(v)
        dst[o] = scale * (-v if e & 1 else v)


def test_dense_permute_tables_with_full_sign_program():
    rng = np.random.RandomState(0)
    shape, stats = (4, 2, 8, 4), (1, -1, -1, 1)
    D = O.random_dense(shape, stats, dtype=complex, rng=rng, skip_trimming=True)
    ref = O.einsum('ijkl->lkij', D).data
    # sign program from the product planner, evaluated per element from popcounts
    from grassmanntn_b200 import _planner as P
    prog, _, _ = P.einsum_sign_program(['ijkl'], 'lkij', [stats])
    labels = 'ijkl'
    n = 4
    in_st = E._row_strides(shape)
    out_shape = tuple(shape[labels.index(c)] for c in 'lkij')
    out_st_by_label = dict(zip('lkij', E._row_strides(out_shape)))
    legs, alpha, beta, Q = [], [], [], [0] * n
    for a, c in enumerate(labels):
        d = shape[a]
        pc = E._popcount_vec(np.arange(d))
        legs.append(E.LegTab(d, np.arange(d) * in_st[a], np.arange(d) * out_st_by_label[c], p=pc & 1, q=(pc >> 1) & 1))
        alpha.append(1 if c in prog.alpha else 0)
        beta.append(1 if c in prog.beta else 0)
    for pr in prog.Q:
        x, y = tuple(pr)
        Q[labels.index(x)] |= 1 << labels.index(y)
        Q[labels.index(y)] |= 1 << labels.index(x)
    fields, tabs = E.build_job(legs, alpha=alpha, beta=beta, Q=Q)
    out = np.zeros(ref.size, dtype=complex)
    emulate(fields, tabs, D.data.ravel(), out)
    assert np.array_equal(out.reshape(ref.shape), ref)
    assert fields["transpose"] == 1 and len(tabs) >= 2


def test_group_layout_parity_sorted():
    lay = E.GroupLayout([(1, 3, 2), (0, 5, 0), (-1, 2, 2)])
    assert lay.total == 5 * 5 * 4
    assert lay.pats[0] == (0, 0) and sum(lay.pats[1]) % 2 == 0
    assert lay.even_total == 3 * 5 * 2 + 2 * 5 * 2
    assert lay.sector(1) == (lay.even_total, lay.total - lay.even_total)
    assert lay.strides((1, 0)) == [10, 2, 1]


def test_choose_groups_fuses_small_axes():
    groups, tr = E._choose_groups([4, 4, 4, 4], [0, 1, 2, 3], [1, 2, 3, 0])
    assert tr and groups[0][-1] == 3 and groups[1][-1] == 0
    sizes = [int(np.prod([4 for _ in g])) for g in groups]
    assert sizes[0] >= 16
    groups, tr = E._choose_groups([8, 8, 64], [0, 1, 2], [1, 0, 2])
    assert not tr and groups[0] == [2]
