"""Inference input loading: `load_pi_instance` (phydata.py:1249-1289) and its helpers.

PHYLIP (sequential or interleaved) and FASTA alignments -> the batch dict the rollout drivers
consume: `data` int8 [1,R,L,4] one-hot (gap / N = 1111, pad '*' = 0000, phydata.py:38-46),
`seq_weights` float32 [1,L], `seqs`, `seq_keys`, ...  Taxa of a .phy file are ordered by the
integer suffix of their names (phydata.py:1252-1262).  The reference's unused O(R^2 L) Hamming
matrix (phydata.py:1273) is not computed.
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

import numpy as np
import torch

_CODE = {"A": (1, 0, 0, 0), "C": (0, 1, 0, 0), "G": (0, 0, 1, 0), "T": (0, 0, 0, 1),
         "-": (1, 1, 1, 1), "N": (1, 1, 1, 1), "*": (0, 0, 0, 0)}
_LUT = np.zeros((256, 4), dtype=np.int8)
_LUT[:] = _CODE["-"]                      # any other symbol is read as a gap (phydata.py:536-543)
for _ch, _v in _CODE.items():
    _LUT[ord(_ch)] = _v


class _Known(dict):
    """str.translate table: the seven symbols of `_CODE` map to themselves, anything else to a gap (phydata.py:536-543)."""

    def __missing__(self, key):
        return "-"


_KNOWN = _Known({ord(c): c for c in _CODE})
_LUT32 = np.ascontiguousarray(_LUT).view(np.uint32).reshape(256)     # one 4-byte gather per symbol instead of four 1-byte ones


def load_phy_file_multirow(path: str) -> Tuple[List[str], List[str], int, int]:
    """PHYLIP reader: first block `name sequence`, further interleaved blocks without names (phydata.py:499-548)."""
    with open(path) as f:
        n_taxa, n_sites = (int(t) for t in f.readline().split())
        lines = [ln.strip() for ln in f]
    names: List[str] = []
    seqs: Dict[str, str] = {}
    k = 0
    while k < len(lines) and lines[k]:
        parts = lines[k].split(maxsplit=1)
        if len(parts) > 1:
            names.append(parts[0])
            seqs[parts[0]] = parts[1].replace(" ", "").upper()
        k += 1
    row = 0
    for ln in lines[k + 1:]:
        if ln:
            seqs[names[row]] += ln.replace(" ", "").upper()
            row += 1
        else:
            row = 0
    if len(names) != n_taxa:
        raise ValueError(f"{path}: header says {n_taxa} taxa, found {len(names)}")
    out = []
    for nm in names:
        if len(seqs[nm]) != n_sites:
            raise ValueError(f"{path}: sequence {nm} has {len(seqs[nm])} sites, header says {n_sites}")
        out.append(seqs[nm].translate(_KNOWN))
    return out, names, n_taxa, n_sites


def load_alignment_file(path: str) -> Tuple[List[str], List[str], int, int]:
    """FASTA / .aln reader (phydata.py `load_alignment_file`): order as in the file."""
    names, seqs = [], []
    with open(path) as f:
        for ln in f:
            ln = ln.strip()
            if not ln:
                continue
            if ln.startswith(">"):
                names.append(ln[1:].split()[0])
                seqs.append("")
            elif names:
                seqs[-1] += ln.replace(" ", "").upper()
    seqs = [s.translate(_KNOWN) for s in seqs]
    return seqs, names, len(names), max((len(s) for s in seqs), default=0)


def encode_sequences(seqs: List[str]) -> Tuple[np.ndarray, np.ndarray]:
    """Pad ragged rows with '*' to the longest row (phydata.py:98-123) and one-hot encode -> int8 [R,L,4], weights [L]."""
    L = max(len(s) for s in seqs)
    cols_real = min(len(s) for s in seqs)
    raw = np.full((len(seqs), L), ord("*"), dtype=np.uint8)
    for r, s in enumerate(seqs):
        raw[r, :len(s)] = np.frombuffer(s.encode("ascii"), dtype=np.uint8)
    weights = np.ones(L, dtype=np.float32)
    # the reference zips the rows into columns (truncating to the shortest row) and pads the rest
    raw[:, cols_real:] = ord("*")
    weights[cols_real:] = 0.0
    return _LUT32[raw].view(np.int8).reshape(raw.shape[0], L, 4), weights


def load_pi_instance(file_path: str) -> dict:
    if file_path.endswith(".phy"):
        seqs, keys, n_taxa, n_sites = load_phy_file_multirow(file_path)
        m = re.match(r"^([a-zA-Z]+)([0-9]+)$", keys[0])
        if m is None:
            raise ValueError(f"{file_path}: taxon names must look like <letters><integer> (got {keys[0]!r})")
        plen = len(m.group(1))
        order = [None] * len(keys)
        for s, k in zip(seqs, keys):
            order[int(k[plen:]) - 1] = (s, k)
        seqs, keys = [p[0] for p in order], [p[1] for p in order]
    elif file_path.endswith(".fasta") or file_path.endswith(".aln"):
        seqs, keys, n_taxa, n_sites = load_alignment_file(file_path)
    else:
        raise ValueError(f"unsupported alignment format: {file_path}")
    data, weights = encode_sequences(seqs)
    return {
        "data": torch.from_numpy(data[None]),
        "seqs": [seqs],
        "seq_keys": [keys],
        "seq_weights": torch.from_numpy(weights[None]),
        "file_paths": [file_path],
        "taxa_nums": [n_taxa],
        "seq_lens": [n_sites],
    }
