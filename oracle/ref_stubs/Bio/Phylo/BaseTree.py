"""Stand-in for Bio.Phylo.BaseTree (Biopython is not installed here): just enough of Clade / Tree for the reference's
phydata.load_tree_file / sorted_bio_tree / sample_trajectory_set_bottom_top to run unmodified on a label tree.
Used by oracle/make_golden.py only."""


class Clade:
    def __init__(self, branch_length=None, name=None, clades=None, confidence=None):
        self.branch_length = branch_length
        self.name = name
        self.clades = list(clades) if clades else []
        self.confidence = confidence

    def is_terminal(self):
        return not self.clades

    def _walk(self):                       # preorder, as Biopython's default traversal
        yield self
        for c in self.clades:
            yield from c._walk()

    def get_terminals(self):
        return [c for c in self._walk() if c.is_terminal()]

    def get_nonterminals(self):
        return [c for c in self._walk() if not c.is_terminal()]


class Tree:
    def __init__(self, root):
        self.root = root
        self.rooted = False

    def get_terminals(self):
        return self.root.get_terminals()

    def get_nonterminals(self):
        return self.root.get_nonterminals()
