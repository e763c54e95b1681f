"""Per-kernel ptxas resource table (registers, spills, static smem) of the shipped build: python scratch/ptxas_summary.py > profiles/r02_ptxas_resources.txt
Runs a forced verbose rebuild (nvcc -Xptxas -v, cross-compiles without a GPU) and demangles the entry names with cu++filt / c++filt."""
import contextlib
import io
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from neuralnj_b200 import _lib

buf = io.StringIO()
with contextlib.redirect_stderr(buf):
    _lib.build(force=True, verbose=True)
rows, cur = [], None
for line in buf.getvalue().splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = {"name": m.group(1), "regs": 0, "spill_st": 0, "spill_ld": 0, "stack": 0, "smem": 0}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur["regs"] == 0:
        cur["stack"], cur["spill_st"], cur["spill_ld"] = map(int, m.groups())
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
        s = re.search(r"(\d+) bytes smem", line)
        cur["smem"] = int(s.group(1)) if s else 0
        cur = None
filt = "cu++filt" if subprocess.run(["which", "cu++filt"], capture_output=True).returncode == 0 else "c++filt"
names = subprocess.run([filt], input="\n".join(r["name"] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"ptxas -v, sm_100a, flags: {' '.join(_lib.NVCC_FLAGS)}; {len(rows)} kernels; dynamic shared memory is set at launch and not listed")
print(f"{'kernel':<64} {'regs':>5} {'spill st/ld B':>14} {'stack B':>8} {'static smem B':>14}")
for r, n in sorted(zip(rows, names), key=lambda x: x[1]):
    n = re.sub(r"\(.*", "", n).replace("nnj::", "").replace("void ", "")
    print(f"{n[:64]:<64} {r['regs']:>5} {str(r['spill_st']) + '/' + str(r['spill_ld']):>14} {r['stack']:>8} {r['smem']:>14}")
