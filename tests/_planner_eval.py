"""Test helper (CPU): evaluate a Grassmann einsum with numpy using ONLY the product's planner
(grassmanntn_b200._planner) for the signs.  Used to check the planner against the oracle
without a GPU."""
import numpy as np

from grassmanntn_b200 import _planner as P


def _ascii(sub):
    table, out = {}, []
    letters = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    for ch in sub:
        if ch in ",->":
            out.append(ch)
        else:
            table.setdefault(ch, letters[len(table)])
            out.append(table[ch])
    return "".join(out)


def planner_einsum(subscripts, arrays, stats, ignore=False):
    inputs, output = P.parse_subscripts(subscripts)
    prog, first_stat, contracted = P.einsum_sign_program(inputs, output, stats, ignore)
    dims = {}
    for sub, arr in zip(inputs, arrays):
        for ch, d in zip(sub, arr.shape):
            dims[ch] = d
    par = lambda d: np.array([bin(i).count("1") & 1 for i in range(d)])
    sig = lambda d: np.array([(bin(i).count("1") >> 1) & 1 for i in range(d)])
    subs, ops = list(inputs), list(arrays)
    for x in prog.alpha:
        subs.append(x); ops.append((-1.0) ** par(dims[x]))
    for x in prog.beta:
        subs.append(x); ops.append((-1.0) ** sig(dims[x]))
    for pr in prog.Q:
        x, y = tuple(pr)
        subs.append(x + y); ops.append((-1.0) ** np.outer(par(dims[x]), par(dims[y])))
    es = ",".join(subs) + ("->" + output if output is not None else "")
    return np.einsum(_ascii(es), *ops, optimize="greedy")
