"""Stub: ete3.Tree is only used for RF distances against label trees (utils.py:255-261)."""


class Tree:
    pass
