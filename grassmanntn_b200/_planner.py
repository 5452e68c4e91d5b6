"""Host-side planner: subscripts -> Grassmann sign program.

Pure Python (no device work), so it is covered by the CPU test-suite against the oracle.

The reference computes the sign of every element of a Grassmann einsum as S1*S2*S3
(reference einsum_ds, __init__.py:1750-2126):
  S1  sign of moving every conjugated contracted index to the immediate right of its
      non-conjugated partner inside the concatenated list of fermionic index occurrences;
  S2  sigma_i = (-1)^{p(p-1)/2} for every contracted fermionic index value i;
  S3  sign of permuting the surviving indices into the requested output order;
where a permutation contributes (-1) for every pair of ODD-parity indices whose order it
flips (absolute_sign / relative_sign, __init__.py:1483-1591).  All three therefore collapse
into ONE quadratic form over GF(2) in the parities p_x of the distinct index labels plus a
linear term in the sigma bits:
      exponent = sum_x alpha_x p_x + sum_{x<y} Q_xy p_x p_y + sum_{x contracted} q_x .
This module derives (alpha, Q, beta) from the subscripts; the kernels evaluate it per element
(dense storage) or per parity block (block storage).
"""
from .param import gparity

CHAR_LIST = tuple("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ") + tuple(
    "αβΓγΔδεζηΘθικλμνΞξΠπρΣσςτυΦϕφχΨψΩω")            # reference __init__.py:32-37
SEPARATORS = ("|", ":", ";", ",", ".")                # reference __init__.py:20
FERMI = (1, -1)
HYBRID = "*"


class GtnValueError(ValueError):
    pass


def get_char(used):
    for ch in CHAR_LIST:
        if ch not in used:
            return ch
    raise GtnValueError("Error[get_char]: Running out of index character!")


def denumerate(string):
    """numbered indices ('i1 i2') -> single characters; same replacement order as the reference
    (__init__.py:1610-1631) so that user-visible index names agree."""
    tokens = []
    for c in string:
        if c.isdigit():
            if not tokens:
                raise GtnValueError("Error[einsum]: subscript cannot start with a digit")
            tokens[-1] += c
        else:
            tokens.append(c)
    multi = sorted([t for t in dict.fromkeys(tokens) if len(t) > 1], key=len, reverse=True)
    pool, repl = string, []
    for t in multi:
        nc = get_char(pool)
        pool += nc
        repl.append((t, nc))
    for t, nc in repl:
        string = string.replace(t, nc)
    return string


def parse_subscripts(subscripts):
    s = denumerate(subscripts.replace(" ", ""))
    if s.count("->") > 1:
        raise GtnValueError("Error[einsum]: Only one arrow is allowed in the input string!")
    if "->" in s:
        lhs, out = s.split("->")
        if "," in out:
            raise GtnValueError("Error[einsum]: Output string must not contain commas (',') !")
    else:
        lhs, out = s, None
    return lhs.split(","), out


class SignProgram:
    """alpha/beta: sets of labels; Q: set of frozenset({x,y}) label pairs."""

    def __init__(self):
        self.alpha = set()
        self.beta = set()
        self.Q = set()

    def flip_pair(self, x, y):
        if x == y:
            self.alpha ^= {x}
        else:
            self.Q ^= {frozenset((x, y))}

    def exponent(self, parity, sigma_bit):
        """evaluate for given label->parity and label->sigma-bit maps (host check / block signs)"""
        e = 0
        for x in self.alpha:
            e ^= parity[x] & 1
        for pr in self.Q:
            x, y = tuple(pr)
            e ^= parity[x] & parity[y] & 1
        for x in self.beta:
            e ^= sigma_bit[x] & 1
        return e

    def block_const(self, parity):
        """alpha and Q part only (block-constant)"""
        e = 0
        for x in self.alpha:
            e ^= parity[x] & 1
        for pr in self.Q:
            x, y = tuple(pr)
            e ^= parity[x] & parity[y] & 1
        return e


def einsum_sign_program(inputs, output, stats_per_operand, ignore_anticommutation=False):
    """Validate the subscripts like reference einsum_ds (:1689-1741) and return
    (SignProgram, label->stat of first occurrence, contracted fermionic labels)."""
    summand = "".join(inputs)
    stats_list = [s for st in stats_per_operand for s in st]
    if len(stats_list) != len(summand):
        raise GtnValueError("Error[einsum]: the number of indices does not match the operands' legs")
    first_stat = {}
    occ = {}                      # label -> list of (position, stat)
    for pos, (ch, st) in enumerate(zip(summand, stats_list)):
        first_stat.setdefault(ch, st)
        occ.setdefault(ch, []).append((pos, st))
    out = output or ""
    contracted = []
    for ch, lst in occ.items():
        st0 = lst[0][1]
        if st0 == 0:
            if any(s != 0 for _, s in lst):
                raise GtnValueError("Error[einsum]: The contracted indices have inconsistent statistics!")
            continue
        if st0 == HYBRID:
            continue
        lf, rf = len(lst), out.count(ch)
        if lf > 2 or rf > 1 or (output is not None and (lf + rf) % 2 == 1):
            raise GtnValueError("Error[einsum]: Inconsistent index statistics.")
        if lf == 2:
            if sorted(s for _, s in lst) != [-1, 1]:
                raise GtnValueError("Error[einsum]: The contracted indices have inconsistent statistics!")
            contracted.append(ch)
    for ch in out:
        if ch not in occ:
            raise GtnValueError("Error[einsum]: output index '%s' does not appear in the inputs" % ch)
    prog = SignProgram()
    if ignore_anticommutation:
        return prog, first_stat, contracted
    # occurrence list of the non-bosonic legs (bosons dropped, :1761); hybrid legs stay like the reference
    seq = [(ch, st) for ch, st in zip(summand, stats_list) if st != 0]
    # ---- S1: conjugated contracted occurrence moves right after its partner
    items = list(range(len(seq)))                    # identify occurrences by position in seq
    partner_pos, conj_pos = {}, {}
    for k, (ch, st) in enumerate(seq):
        if ch in contracted:
            (conj_pos if st == -1 else partner_pos)[ch] = k
    moved = [k for k in items if k not in conj_pos.values()]
    order_after = []
    for k in moved:
        order_after.append(k)
        ch = seq[k][0]
        if ch in contracted and partner_pos.get(ch) == k:
            order_after.append(conj_pos[ch])
    rank = {k: r for r, k in enumerate(order_after)}
    for a in range(len(items)):
        for b in range(a + 1, len(items)):
            if rank[a] > rank[b]:
                prog.flip_pair(seq[a][0], seq[b][0])
    # ---- S2
    for ch in contracted:
        prog.beta ^= {ch}
    # ---- S3: survivors (in order_after order) -> output order
    if output is not None:
        survivors = [seq[k][0] for k in order_after if seq[k][0] not in contracted]
        fout = [ch for ch in out if first_stat[ch] != 0]
        if sorted(survivors) != sorted(fout):
            raise GtnValueError("Error[einsum]: Inconsistent index statistics.")
        pos = {ch: r for r, ch in enumerate(fout)}
        for a in range(len(survivors)):
            for b in range(a + 1, len(survivors)):
                if pos[survivors[a]] > pos[survivors[b]]:
                    prog.flip_pair(survivors[a], survivors[b])
    return prog, first_stat, contracted


def split_partition(string, who="svd"):
    """'ab|cd' or '(ab)(cd)' -> ('ab','cd'); reference svd (:4046-4099)."""
    string = denumerate(string.replace(" ", ""))
    if string.count("(") == string.count(")") and string.count("(") > 0:
        string = string.replace(")(", "|")
        if string.count("(") > 1 or string.count(")") < 1:
            raise GtnValueError("Error[%s]: Parentheses don't match" % who)
        string = string.replace(")", "").replace("(", "")
    if sum(string.count(s) for s in SEPARATORS) != 1:
        raise GtnValueError("Error[%s]: The input string must contain one and only one partition "
                            "( '|', ':', ';', ',', or '.' ) in it." % who)
    for s in SEPARATORS:
        if s in string:
            left, right = string.split(s)
            return left, right
    raise AssertionError


def parse_groups(grouping_string):
    """'(ab)(cd)e' / 'ab|cd' -> list of index strings; reference get_group_info (:3172-3217)."""
    s = grouping_string
    if "(" not in s and ")" not in s:
        for sep in SEPARATORS:
            s = s.replace(sep, ")(")
        s = "(" + s + ")"
    elif any(sep in s for sep in SEPARATORS):
        raise GtnValueError("Error[get_grouping_info]: Do not mix the string format.")
    groups, outside = [], True
    for ch in s:
        if ch == "(" and outside:
            outside = False
            groups.append("")
        elif ch == ")" and not outside:
            outside = True
        elif ch in "()":
            raise GtnValueError("Error[get_grouping_info]: No nested parenthesis allowed!")
        elif outside:
            groups.append(ch)
        else:
            groups[-1] += ch
    return groups


def sigma_bit(i):
    return (gparity(i) >> 1) & 1
