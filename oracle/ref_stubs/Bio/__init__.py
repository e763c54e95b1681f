"""Stub: Bio.Phylo is only used by training-set loaders (phydata.py:7-8)."""
from . import Phylo
