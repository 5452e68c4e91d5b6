"""CPU: the truncated SVD's iteration memory (_engine._trunc_accept) -- pure host bookkeeping.
A count is kept per (batch shape, call site); margins lower it, clean accepts never raise it, after PROBE_EVERY
accepts of the same count it is dropped so that the next run derives it afresh, and fruitless probes become rarer."""
import pytest


@pytest.fixture
def E():
    from grassmanntn_b200 import _engine as E
    saved = (dict(E._trunc_iters_hint), dict(E._trunc_rate), dict(E._trunc_fail), dict(E._trunc_probe), E.ADAPT[0])
    for d in (E._trunc_iters_hint, E._trunc_rate, E._trunc_fail, E._trunc_probe):
        d.clear()
    E.ADAPT[0] = True
    yield E
    for d, s in zip((E._trunc_iters_hint, E._trunc_rate, E._trunc_fail, E._trunc_probe), saved[:4]):
        d.clear(); d.update(s)
    E.ADAPT[0] = saved[4]


K = ("shape", "site")


def test_margin_lowers_the_count_but_keeps_spare_factors(E):
    E._trunc_rate[K] = 0.1
    E._trunc_accept(K, 6, 1e-15, spare=1)              # four factors of 0.1 below 1e-11: three fewer next time
    assert E._trunc_iters_hint[K] == 3
    E._trunc_iters_hint.clear()
    E._trunc_accept(K, 6, 1e-15, spare=2)              # a speculative run keeps two spare factors
    assert E._trunc_iters_hint[K] == 4
    E._trunc_iters_hint.clear()
    E._trunc_accept(K, 3, 5e-12, spare=1)              # passed without margin: unchanged
    assert E._trunc_iters_hint[K] == 3
    E._trunc_iters_hint.clear()
    E._trunc_accept(K, 2, 1e-16, spare=0)              # never below half
    assert E._trunc_iters_hint[K] == 1


def test_clean_accept_never_raises_and_unclean_does(E):
    E._trunc_iters_hint[K] = 2
    E._trunc_accept(K, 5, 5e-12, clean=True)           # a recorded graph still replaying 5 iterations
    assert E._trunc_iters_hint[K] == 2
    E._trunc_accept(K, 5, 5e-12, clean=False)          # the first check failed: 5 are needed
    assert E._trunc_iters_hint[K] == 5 and E._trunc_fail[K] == 0


def test_probe_after_stable_accepts_and_backoff(E):
    E._trunc_accept(K, 4, 5e-12)
    assert E._trunc_iters_hint[K] == 4
    for _ in range(E.PROBE_EVERY - 1):
        E._trunc_accept(K, 4, 5e-12)
        assert E._trunc_iters_hint[K] == 4
    E._trunc_accept(K, 4, 5e-12)                       # PROBE_EVERY accepts of the same count: derive it afresh
    assert K not in E._trunc_iters_hint and E._trunc_probe[K]["before"] == 4
    E._trunc_accept(K, 1, 5e-12)                       # the fresh run needed 1: it paid, interval unchanged
    assert E._trunc_iters_hint[K] == 1 and E._trunc_probe[K]["every"] == E.PROBE_EVERY
    for _ in range(E.PROBE_EVERY):
        E._trunc_accept(K, 1, 5e-12)
    assert K not in E._trunc_iters_hint
    E._trunc_accept(K, 1, 5e-12)                       # nothing gained this time: probe four times more rarely
    assert E._trunc_iters_hint[K] == 1 and E._trunc_probe[K]["every"] == 4 * E.PROBE_EVERY
    for _ in range(4 * E.PROBE_EVERY - 1):
        E._trunc_accept(K, 1, 5e-12)
        assert E._trunc_iters_hint[K] == 1
    for _ in range(3 * E.PROBE_EVERY):
        E._trunc_accept(K, 0, 5e-12, clean=True)       # count 0 cannot go lower: never probed
    assert E._trunc_iters_hint[K] == 0


def test_stale_graph_replays_do_not_count_as_stable(E):
    E._trunc_iters_hint[K] = 2
    for _ in range(3 * E.PROBE_EVERY):
        E._trunc_accept(K, 5, 5e-12, clean=True)       # replays at a recorded count of 5 say nothing about 2
    assert E._trunc_iters_hint[K] == 2


def test_frozen(E):
    E._trunc_iters_hint[K] = 4
    E._trunc_rate[K] = 0.1
    E.ADAPT[0] = False
    for _ in range(3 * E.PROBE_EVERY):
        E._trunc_accept(K, 4, 1e-15)
    assert E._trunc_iters_hint[K] == 4                 # no lowering, no probing
    E._trunc_accept(K, 6, 5e-12, clean=False)          # but a failed first check still teaches the needed count
    assert E._trunc_iters_hint[K] == 6
    E._trunc_accept(("other", "site"), 3, 5e-12)       # and an unknown site still gets its first count
    assert E._trunc_iters_hint[("other", "site")] == 3
