"""Per-kernel SASS opcode summary of neuralnj_b200/libnnj.so -> profiles/r02_sass_opcodes.txt
(tcgen05 = UTC*MMA, tensor-memory loads / stores = LDTM / STTM, TMA = UTMALDG / UTMASTG / UBLKCP, legacy tensor path = HMMA, MUFU)."""
import collections, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
so = os.path.join(root, "neuralnj_b200", "libnnj.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UTMAPF", r"\bUTMAPF"),
        ("HMMA(mma.sync)", r"\bHMMA"), ("LDSM", r"\bLDSM"), ("MUFU", r"\bMUFU"), ("FFMA2/FADD2/FMUL2", r"\bF(FMA|ADD|MUL)2"), ("DFMA", r"\bDFMA")]
cur, counts, order = None, collections.defaultdict(collections.Counter), []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("nnj::", "").replace("void ", "")
        order.append(cur)
        continue
    if cur and "/*" in line:
        for name, p in pats:
            if re.search(p, line):
                counts[cur][name] += 1
                break
        counts[cur]["instructions"] += 1 if re.match(r"\s+/\*[0-9a-f]{4}\*/", line) else 0
out = [f"cuobjdump -sass neuralnj_b200/libnnj.so (sm_100a), instruction counts per kernel; built from commit-tree sources by __graft_entry__.build()",
       f"{'kernel':52s} {'instr':>7s} " + " ".join(f"{n:>9s}" for n, _ in pats)]
tot = collections.Counter()
for k in sorted(set(order), key=lambda k: (-counts[k]["UTC*MMA"] * 1000 - counts[k]["HMMA(mma.sync)"], k)):
    c = counts[k]
    out.append(f"{k[:52]:52s} {c['instructions']:7d} " + " ".join(f"{c[n]:9d}" for n, _ in pats))
    tot.update(c)
out.append(f"{'total':52s} {tot['instructions']:7d} " + " ".join(f"{tot[n]:9d}" for n, _ in pats))
open(os.path.join(root, "profiles", "r02_sass_opcodes.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]), "\n...", out[-1])
