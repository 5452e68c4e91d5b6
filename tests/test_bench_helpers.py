"""CPU: bench.py's host-side helpers (no GPU work)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class FakeG:
    def __init__(self):
        self.STEP_GRAPH_STATS = {"captured": 0, "replayed": 0, "failed_capture": 0, "failed_certificate": 0}


def test_settle_waits_for_quiet_replays():
    import bench
    g = FakeG()
    script = iter(["eager", "eager", "capture", "replay", "capture", "replay", "replay", "replay", "replay"])
    calls = []

    def fn():
        kind = next(script)
        calls.append(kind)
        if kind == "capture":
            g.STEP_GRAPH_STATS["captured"] += 1
            g.STEP_GRAPH_STATS["replayed"] += 1          # a recording call also replays
        elif kind == "replay":
            g.STEP_GRAPH_STATS["replayed"] += 1
    n = bench.settle(g, fn)
    assert n == 7 and calls[-2:] == ["replay", "replay"] and calls[-3] == "replay" or calls[-3] == "capture"
    assert calls == ["eager", "eager", "capture", "replay", "capture", "replay", "replay"]


def test_settle_gives_up_for_graphless_workloads():
    import bench
    g = FakeG()
    count = [0]

    def fn():
        count[0] += 1
    assert bench.settle(g, fn, max_steps=5) == 5 and count[0] == 5


def test_config_and_parse_defaults(monkeypatch):
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    args = bench.parse()
    assert (args.gpus, args.chi, args.impl) == (1, 32, "ours") and args.warmup >= 3 and args.steps >= 1
    cfg = bench.config_dict(args, "label", (32, 32, 32, 32))
    assert "workload" in cfg and "batch" in cfg and "l2" in cfg and "model" not in cfg
