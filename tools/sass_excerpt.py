"""SASS excerpts of the shipped library for profiles/: the TMA staging (UBLKCP + mbarrier) and
the BVH node loops of the batched frame kernel k_frame<true, true, true, true>. CPU only:

    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.md
"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "rvpt_b200" / "librvpt_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
# split per function
funcs, name, cur = {}, None, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name:
            funcs[name] = cur
        name, cur = m.group(1), []
    elif re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        cur.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())
if name:
    funcs[name] = cur
key = next(k for k in funcs if "k_frameILb1ELb1ELb1ELb1E" in k)
ins = funcs[key]
print(f"# SASS excerpts of `{LIB.name}` (cuobjdump -sass), kernel `k_frame<true, true, true, true>` — "
      f"{len(ins)} instructions, sm_100a\n")
print("All frame-kernel instantiations in the library: " + ", ".join(
    sorted(re.search(r"k_frameILb(\d)ELb(\d)ELb(\d)ELb(\d)E", k).group(0) for k in funcs if "k_frameILb" in k)) + "\n")

def show(title, lo, hi):
    print(f"## {title}\n\n```")
    print("\n".join(ins[max(lo, 0):hi]))
    print("```\n")

i = next(k for k, l in enumerate(ins) if "UBLKCP" in l)
show("Scene staging: mbarrier init / expect-tx, TMA bulk copies global -> shared (`cp.async.bulk`), try-wait", i - 14, i + 22)
# node loops: windows with two LDS.128, FMNMX and FMNMX3 close together
found, k = [], 0
while k < len(ins) and len(found) < 2:
    w = ins[k:k + 30]
    if sum("LDS.128" in l for l in w[:6]) >= 2 and sum("FMNMX3" in l for l in w) >= 2 and "0x19000" in " ".join(w[:6]):
        found.append(k)
        k += 60
    else:
        k += 1
titles = ["Node loop, primary rays (octant arrays relative to the camera origin: 6 FMUL, no FADD)",
          "Node loop, bounce rays (6 FADD + 6 FMUL)"]
for t, k in zip(titles, found):
    show(t, k - 3, k + 30)
print("Counts over the whole kernel: " + ", ".join(
    f"{op} {sum(op in l for l in ins)}" for op in ("UBLKCP", "SYNCS", "LDS.128", "FMNMX3", "MATCH.ANY", "VOTE", "REDUX", "ATOMG", "RED.", "STL", "LDL", "FFMA")))
