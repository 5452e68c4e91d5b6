"""CPU: the product's host logic (planner, table builder) against the oracle, no GPU needed."""
import itertools

import numpy as np
import pytest

import gtn_oracle as O
from _planner_eval import planner_einsum
from grassmanntn_b200 import _planner as P

CASES = [
    ('ijkl->jkli', [((4, 4, 4, 4), (1, 1, -1, -1))]),
    ('ijkl->klij', [((4, 2, 8, 4), (1, -1, -1, 1))]),
    ('ijkl,klmn->ijmn', [((4, 4, 4, 2), (1, 1, 1, 1)), ((4, 2, 4, 8), (-1, -1, 1, -1))]),
    ('kwz,lxw->lxzk', [((4, 4, 8), (1, -1, 1)), ((4, 2, 4), (-1, 1, 1))]),
    ('lxzk,jzxi->ijkl', [((4, 2, 8, 4), (1, 1, -1, 1)), ((2, 8, 2, 4), (-1, 1, -1, 1))]),
    ('ijij', [((4, 2, 4, 2), (1, 1, -1, -1))]),
    ('ijij', [((4, 2, 4, 2), (-1, 1, 1, -1))]),
    ('ijkl,klij', [((4, 2, 4, 8), (1, 1, -1, -1)), ((4, 8, 4, 2), (1, 1, -1, -1))]),
    ('IJIK,iKiJ', [((4, 2, 4, 8), (1, 1, -1, -1)), ((2, 8, 2, 2), (1, 1, -1, -1))]),
    ('IJKLij,ij->IJKL', [((4, 2, 4, 2, 3, 3), (1, 1, -1, -1, 0, 0)), ((3, 3), (0, 0))]),
    ('i1 i3 a, j1 j3 b -> i1 i3 ab j1 j3', [((4, 2, 3), (1, -1, 0)), ((2, 4, 2), (1, -1, 0))]),
    ('abx,xc->abc', [((4, 4, 8), (1, 1, 1)), ((8, 8), (-1, 1))]),
    ('ab,bc,cd->ad', [((4, 8), (1, 1)), ((8, 2), (-1, 1)), ((2, 4), (-1, -1))]),
    ('abc,dbe,fce->adf', [((4, 8, 2), (1, 1, 1)), ((2, 8, 4), (1, -1, 1)), ((4, 2, 4), (-1, -1, -1))]),
    ('ajk,jib->aibk', [((4, 2, 8), (-1, -1, -1)), ((2, 4, 4), (1, 1, 1))]),
    ('ax,bx->ba', [((2, 4), (1, -1)), ((8, 4), (-1, 1))]),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_sign_program_matches_oracle(case):
    sub, ops = CASES[case]
    rng = np.random.RandomState(case)
    Ds = [O.random_dense(s, st, dtype=complex, rng=rng, skip_trimming=True) for s, st in ops]
    ref = O.einsum(sub, *Ds)
    got = planner_einsum(sub, [d.data for d in Ds], [d.statistics for d in Ds])
    ref = ref.data if isinstance(ref, O.Dense) else ref
    assert np.abs(np.asarray(ref) - got).max() <= 1e-12 * max(np.abs(np.asarray(ref)).max(), 1)


def test_random_single_operand_permutations_bit_exact():
    rng = np.random.RandomState(3)
    letters = "abcde"
    for trial in range(40):
        n = rng.randint(2, 6)
        dims = [int(2 ** rng.randint(0, 3)) for _ in range(n)]
        stats = [int(rng.choice([1, -1, 0])) for _ in range(n)]
        perm = rng.permutation(n)
        sub = letters[:n] + "->" + "".join(letters[p] for p in perm)
        D = O.random_dense(dims, stats, dtype=float, rng=rng, skip_trimming=True)
        ref = O.einsum(sub, D).data
        got = planner_einsum(sub, [D.data], [D.statistics])
        assert np.array_equal(ref, got), sub


def test_ignore_anticommutation():
    rng = np.random.RandomState(4)
    A = O.random_dense((4, 4, 4), (1, -1, 1), dtype=float, rng=rng, skip_trimming=True)
    B = O.random_dense((4, 4, 2), (-1, 1, 1), dtype=float, rng=rng, skip_trimming=True)
    ref = O.einsum('abc,cbd->ad', A, B, ignore_anticommutation=True).data
    assert np.allclose(ref, np.einsum('abc,cbd->ad', A.data, B.data))
    got = planner_einsum('abc,cbd->ad', [A.data, B.data], [A.statistics, B.statistics], ignore=True)
    assert np.allclose(got, ref)


def test_error_behaviour():
    with pytest.raises(ValueError):
        P.parse_subscripts("ab->c->d")
    with pytest.raises(ValueError):
        P.einsum_sign_program(["ab", "bc"], "ac", [(1, 1), (1, 1)])      # contracted pair not (+1,-1)
    with pytest.raises(ValueError):
        P.einsum_sign_program(["aab"], "ab", [(1, -1, 1)])               # kept fermionic index repeated
    with pytest.raises(ValueError):
        P.split_partition("abcd", "svd")
    assert P.split_partition("(ab)(cd)") == ("ab", "cd")
    assert P.split_partition("i1 i2 | j1") [1] != ""
    assert P.parse_groups("(ab)(cd)e") == ["ab", "cd", "e"]


def test_denumerate_matches_oracle():
    for s in ["i1i2i3->i3i1i2", "a10a1b->ba1a10", "I1 J1 i3 j3 ab".replace(" ", "")]:
        assert P.denumerate(s) == O.denumerate(s)
