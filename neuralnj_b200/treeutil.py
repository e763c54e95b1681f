"""Newick helpers used on the Argmax path: `treestr_to_tuples` (RAxMLpy/raxmlpy/core.py:24-29, the one
pure-Python raxmlpy function Argmax touches via environment.py:680) and an ete3-free RF distance."""
from __future__ import annotations

from typing import Set


def treestr_to_tuples(treestr: str):
    """'(a:0.1, (b:0.2, c:0.3):0.4);' -> ('a', 0.1, ('b', 0.2, 'c', 0.3), 0.4)  (child, length, child, length...)."""
    s = treestr.strip().rstrip(";")
    pos = 0

    def skip_ws():
        nonlocal pos
        while pos < len(s) and s[pos] == " ":
            pos += 1

    def node():
        nonlocal pos
        skip_ws()
        if s[pos] == "(":
            pos += 1
            items = []
            while True:
                child = node()
                skip_ws()
                length = None
                if pos < len(s) and s[pos] == ":":
                    pos += 1
                    st = pos
                    while pos < len(s) and s[pos] not in ",)":
                        pos += 1
                    length = float(s[st:pos])
                items.append(child)
                if length is not None:
                    items.append(length)
                skip_ws()
                if s[pos] == ",":
                    pos += 1
                    continue
                if s[pos] == ")":
                    pos += 1
                    break
                raise ValueError(f"bad Newick near position {pos}")
            return tuple(items)
        st = pos
        while pos < len(s) and s[pos] not in ":,)":
            pos += 1
        return s[st:pos].strip()

    return node()


def bipartitions(newick: str) -> Set[frozenset]:
    """Non-trivial leaf bipartitions of an unrooted reading of `newick`."""
    tup = treestr_to_tuples(newick)
    splits = []

    def walk(t):
        if isinstance(t, str):
            return frozenset([t])
        leaves = frozenset()
        for item in t:
            if isinstance(item, (tuple, str)):
                leaves |= walk(item)
        splits.append(leaves)
        return leaves

    everything = walk(tup)
    out = set()
    for sp in splits:
        if 1 < len(sp) < len(everything) - 1:
            other = everything - sp
            out.add(min(sp, other, key=lambda t: (len(t), sorted(t))))
    return out


def rf_distance(newick_a: str, newick_b: str) -> int:
    """Unrooted Robinson-Foulds distance (what `t1.compare(t2, unrooted=True)['rf']` returns, utils.py:255-261)."""
    return len(bipartitions(newick_a) ^ bipartitions(newick_b))


def normalized_rf(newick_a: str, newick_b: str) -> float:
    a, b = bipartitions(newick_a), bipartitions(newick_b)
    tot = len(a) + len(b)
    return len(a ^ b) / tot if tot else 0.0
