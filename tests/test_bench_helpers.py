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
    assert (args.gpus, args.chi, args.impl) == (1, 128, "ours") and args.warmup >= 3 and args.steps >= 1
    cfg = bench.config_dict(args, "label", "replicas x1")
    assert "workload" in cfg and "batch" in cfg and "l2" in cfg and "model" not in cfg
    assert "chi=128" in cfg["workload"] and "128x128x128x128" in cfg["workload"]


def test_cpu_sample_levels():
    """the CPU arm runs full steps up to chi = 64 and scales the D = 64 step by (chi/64)^6 above"""
    import bench
    assert bench.cpu_level(32) == (32, 1.0) and bench.cpu_level(64) == (64, 1.0)
    lvl, scale = bench.cpu_level(128)
    assert lvl == 64 and scale == 64.0


def test_cpu_reference_arm_small_chi(monkeypatch):
    """--impl reference at a small chi: full oracle steps, a steps/s value and a sample description"""
    import argparse
    import bench
    data, stats, _ = bench.load_z2()
    r = bench.cpu_reference_run(argparse.Namespace(chi=8, steps=1, warmup=0), data, stats)
    assert r["value"] > 0 and "full TRG steps" in r["sample"]


def test_leg_ranges_partition():
    from grassmanntn_b200 import sharded
    for n in (0, 1, 7, 8, 64, 65):
        for w in (1, 2, 3, 8):
            rs = [sharded.leg_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
