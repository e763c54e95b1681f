"""Stub for the reference's native `raxmlpy` package (RAxML-NG binding: not buildable offline).

Only used by oracle/make_golden.py inside the build container.  The two pure-Python
helpers the Argmax path touches (environment.py:680) are taken from the reference's
own RAxMLpy/raxmlpy/core.py at import time; the native entry points raise.
"""
import os as _os

_core = _os.path.join(_os.environ.get("NNJ_REFERENCE", "/root/reference"), "RAxMLpy", "raxmlpy", "core.py")
_src = "\n".join(l for l in open(_core).read().split("\n") if "cpp_binding" not in l)
exec(compile(_src, _core, "exec"), globals())


def optimize_brlen(*a, **k):
    raise RuntimeError("raxmlpy native binding is not available in this container")


def compute_llh(*a, **k):
    raise RuntimeError("raxmlpy native binding is not available in this container")
