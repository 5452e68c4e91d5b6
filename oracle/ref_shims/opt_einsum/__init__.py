"""TEST INFRASTRUCTURE ONLY -- stand-in for `opt_einsum`.

`contract(subscripts, *operands)` -> numpy.einsum(..., optimize='greedy') after
remapping the subscripts to ASCII letters (the reference uses Greek letters,
numpy only accepts a-zA-Z).  COO-shim operands are densified; if any operand was
a COO-shim the result is wrapped again.  Used only by oracle/ref_harness.py.
"""
import string

import numpy as np

_ASCII = string.ascii_letters


def contract(subscripts, *operands, **_kw):
    try:
        from sparse import COO
    except Exception:  # pragma: no cover
        COO = ()
    was_sparse = any(isinstance(o, COO) for o in operands) if COO else False
    ops = [o.todense() if (COO and isinstance(o, COO)) else np.asarray(o) for o in operands]
    table = {}
    out = []
    for ch in subscripts:
        if ch in ",->":
            out.append(ch)
            continue
        if ch not in table:
            if len(table) >= len(_ASCII):
                raise ValueError("opt_einsum shim: more than 52 distinct indices")
            table[ch] = _ASCII[len(table)]
        out.append(table[ch])
    res = np.einsum("".join(out), *ops, optimize="greedy")
    if was_sparse:
        if np.ndim(res) > 0:
            return COO.from_numpy(res)
        out = COO.__new__(COO)                  # 0-d result: one explicit entry, like the real package
        out.shape = ()
        out.coords = np.zeros((0, 1), dtype=np.int64)
        out.data = np.array([res])
        out._dtype = np.asarray(res).dtype
        return out
    return res
