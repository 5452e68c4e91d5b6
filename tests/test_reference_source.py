"""Drop-in proof: the reference's OWN coarse-graining source -- the text of `trg` in
/root/reference/gauge2d.py (:1647-1755) and gauge2d_block.py (:1649-1755) -- is executed UNMODIFIED with `gtn` bound to
this package, on the CPU through the host test double, and must reproduce the real reference's golden numbers.
Skipped where /root/reference does not exist (the GPU box).  Also: `import grassmanntn` resolves to this package when
the repository root is on sys.path."""
import ast
import os
import subprocess
import sys
import time

import numpy as np
import pytest

import _host_double

REF = "/root/reference"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def gtn_host(monkeypatch):
    gtn, saved = _host_double.install(monkeypatch)
    yield gtn
    _host_double.uninstall(saved)


def _reference_function(path, name, gtn):
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    text = ast.get_source_segment(src, node)
    ns = {"gtn": gtn, "np": np, "time": time}
    exec(compile(text, path, "exec"), ns)
    return ns[name], text


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
@pytest.mark.parametrize("fmt,module,key,cut", [("dense", "gauge2d.py", "dense_trg_chi16", 16),
                                                  ("block", "gauge2d_block.py", "block_trg_chi32", 32)])
def test_reference_trg_source_unmodified(gtn_host, fmt, module, key, cut):
    gtn = gtn_host
    trg_ref, text = _reference_function(os.path.join(REF, module), "trg", gtn)
    assert "gtn.einsum('lxzk,jzxi->ijkl',VV,UU)" in text            # it really is the reference's text
    ref = np.load(os.path.join(G, "z2_cg.npz"))[key]
    g = gtn.gauge2d
    T = g.zcap(g.load_initial_tensor())
    if fmt == "block":
        T = T.toblock()
    logNorm = 0.0
    for i in range(2):
        T, Tn, err = trg_ref(T, cut, iternum=i, error_test=True)
        logNorm = 2 * logNorm + np.log(Tn)
        F = (g.logZ(T, "anti-periodic") + logNorm) / 2 ** (i + 1)
        assert abs(Tn - ref[i + 1, 0]) <= 1e-10 * ref[i + 1, 0], (i, Tn, ref[i + 1, 0])
        assert abs(F - complex(ref[i + 1, 2], ref[i + 1, 3])) <= 1e-10 * abs(F), (i, F)
        assert abs(err - ref[i + 1, 1]) <= 1e-8 * max(ref[i + 1, 1], 1e-3)
        assert isinstance(T, gtn.block if fmt == "block" else gtn.dense)


def test_import_alias():
    """`import grassmanntn as gtn` (the reference's package name) gives the B200 package; run in a fresh interpreter so
    that the alias does not shadow the real reference that other tests import under the same name"""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import grassmanntn as gtn\n"
            "from grassmanntn import param\n"
            "from grassmanntn import gauge2d_block as gauge\n"
            "import grassmanntn_b200\n"
            "assert gtn is grassmanntn_b200 and param is grassmanntn_b200.param and gauge is grassmanntn_b200.gauge2d\n"
            "assert param.gparity(7) == 3 and hasattr(gtn, 'einsum') and hasattr(gauge, 'trg')\n"
            "print('alias ok')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "alias ok" in out.stdout, out.stderr[-2000:]
