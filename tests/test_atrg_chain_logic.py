"""CPU: control flow of the speculative ATRG step (gauge2d.atrg2dy) with the ops replaced by symbolic stand-ins.
The three decompositions are enqueued without waiting for each other's certificate; afterwards the certificates are
verified in order, the first failed stage is resumed, later (unverified-input) stages are discarded and redone."""
import pytest


class FakePending:
    def __init__(self, site, ok, log):
        self.site, self._ok, self.log, self.ok = site, ok, log, None

    def verify(self):
        self.log.append(("verify", self.site[2], self._ok))
        self.ok = self._ok
        return self._ok

    def discard(self):
        self.log.append(("discard", self.site[2]))
        self.ok = False


def _run(monkeypatch, fail_plan, speculate=True):
    """fail_plan: {stage index (1..3): versions (1 = first run) of that stage whose certificate fails}"""
    import grassmanntn_b200 as gtn
    from grassmanntn_b200 import gauge2d as g
    log, fails, version = [], dict(fail_plan), {}

    def einsum(sub, *ops):
        return ("E", sub.replace(" ", "")) + tuple(ops)

    def sqrt(x):
        return ("sqrt", x)

    def svd_many(objs, string, cutoff=None, speculative=False, resume=None, site=None):
        stage = site[2]
        version[stage] = version.get(stage, 0) + 1
        tag = (stage, version[stage], "resumed" if resume is not None else ("spec" if speculative else "plain"))
        log.append(("svd",) + tag + (objs[0],))
        res = [(("U",) + tag + (o,), ("S",) + tag + (o,), ("V",) + tag + (o,)) for o in objs]
        if resume is not None:
            assert resume.site == site and resume.ok is False
            return res
        if speculative:
            ok = version[stage] not in fails.get(stage, ())
            return res, FakePending(site, ok, log)
        return res
    monkeypatch.setattr(gtn, "einsum", einsum)
    monkeypatch.setattr(gtn, "sqrt", sqrt)
    monkeypatch.setattr(gtn, "svd_many", svd_many)
    monkeypatch.setattr(g, "_normalised", lambda T: (T, 1.0))
    monkeypatch.setattr(g, "SPECULATE", speculate)
    monkeypatch.setattr(g, "STEP_GRAPH", False)          # the eager chain is under test here
    g.SPEC_STATS["speculated"] = g.SPEC_STATS["failed"] = 0
    T = ("T0",)
    out, norm = g.atrg2dy(T, T, 32)
    return out, log, dict(g.SPEC_STATS)


def _used_svds(expr, acc=None):
    """(stage, version, mode) of every decomposition result the final tensor was built from"""
    acc = set() if acc is None else acc
    if isinstance(expr, tuple):
        if expr and expr[0] in ("U", "S", "V"):
            acc.add(expr[1:4])
            _used_svds(expr[4], acc)            # ... and from whatever that decomposition's input was built from
        else:
            for x in expr:
                _used_svds(x, acc)
    return acc


def test_all_certificates_pass(monkeypatch):
    out, log, st = _run(monkeypatch, {})
    assert [x[:4] for x in log if x[0] == "svd"] == [("svd", 1, 1, "spec"), ("svd", 2, 1, "spec"), ("svd", 3, 1, "spec")]
    assert [x for x in log if x[0] != "svd"] == [("verify", 1, True), ("verify", 2, True), ("verify", 3, True)]
    assert _used_svds(out) == {(1, 1, "spec"), (2, 1, "spec"), (3, 1, "spec")}
    assert st == {"speculated": 3, "failed": 0}


def test_middle_stage_fails_is_resumed_and_tail_redone(monkeypatch):
    out, log, st = _run(monkeypatch, {2: {1}})
    svds = [x[1:4] for x in log if x[0] == "svd"]
    assert svds == [(1, 1, "spec"), (2, 1, "spec"), (3, 1, "spec"), (2, 2, "resumed"), (3, 2, "spec")]
    other = [x for x in log if x[0] != "svd"]
    assert other == [("verify", 1, True), ("verify", 2, False), ("discard", 3), ("verify", 3, True)]
    # the result rests on the verified stage 1, the resumed stage 2 and the repeated stage 3 only
    assert _used_svds(out) == {(1, 1, "spec"), (2, 2, "resumed"), (3, 2, "spec")}
    # the resumed decomposition got the same input as the failed one
    ins = {x[1:3]: x[4] for x in log if x[0] == "svd"}
    assert ins[(2, 1)] == ins[(2, 2)]
    assert st["failed"] == 1


def test_every_stage_fails_once(monkeypatch):
    out, log, st = _run(monkeypatch, {1: {1}, 2: {2}, 3: {3}})
    svds = [x[1:4] for x in log if x[0] == "svd"]
    assert svds == [(1, 1, "spec"), (2, 1, "spec"), (3, 1, "spec"),
                    (1, 2, "resumed"), (2, 2, "spec"), (3, 2, "spec"),
                    (2, 3, "resumed"), (3, 3, "spec"),
                    (3, 4, "resumed")]
    assert _used_svds(out) == {(1, 2, "resumed"), (2, 3, "resumed"), (3, 4, "resumed")}
    assert st["failed"] == 3


def test_not_speculating(monkeypatch):
    out, log, st = _run(monkeypatch, {}, speculate=False)
    assert [x[1:4] for x in log if x[0] == "svd"] == [(1, 1, "plain"), (2, 1, "plain"), (3, 1, "plain")]
    assert st == {"speculated": 0, "failed": 0}
