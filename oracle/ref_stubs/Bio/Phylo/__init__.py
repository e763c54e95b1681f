from . import BaseTree
from .BaseTree import Clade, Tree


def read(handle, fmt):
    """Minimal Newick reader: names, branch lengths, nesting (no quoting, no comments)."""
    assert fmt == "newick"
    s = handle.read().strip()
    pos = 0

    def label(node):
        nonlocal pos
        start = pos
        while pos < len(s) and s[pos] not in ",():;":
            pos += 1
        txt = s[start:pos].strip()
        if txt:
            node.name = txt
        if pos < len(s) and s[pos] == ":":
            pos += 1
            start = pos
            while pos < len(s) and s[pos] not in ",();":
                pos += 1
            node.branch_length = float(s[start:pos])

    def clade():
        nonlocal pos
        node = Clade()
        if s[pos] == "(":
            pos += 1
            while True:
                node.clades.append(clade())
                if s[pos] == ",":
                    pos += 1
                    continue
                assert s[pos] == ")", (pos, s[pos])
                pos += 1
                break
        label(node)
        return node

    return Tree(clade())
