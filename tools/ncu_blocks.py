"""Summarise an `ncu --page source --csv` dump into basic-block-like groups:
contiguous SASS instructions with the same execution count.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_blocks.py src.csv [min_share%]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
# first kernel section only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    if len(r) > 10:
        data.append(r)
iS, iE, iN, iT = (hdr.index(k) for k in ("Source", "Instructions Executed", "# Samples",
                                          "Avg. Threads Executed"))
tot = sum(int(r[iE]) for r in data)
totS = sum(int(r[iN]) for r in data) or 1
print(f"{len(data)} SASS instructions, {tot} warp-instructions executed, {totS} samples")
blocks, cur = [], None
for k, r in enumerate(data):
    e = int(r[iE])
    if cur and abs(e - cur["e"]) <= 0.02 * max(e, cur["e"], 1):
        cur["n"] += 1; cur["inst"] += e; cur["smp"] += int(r[iN]); cur["end"] = k
    else:
        cur = {"start": k, "end": k, "e": e, "n": 1, "inst": e, "smp": int(r[iN]), "thr": r[iT]}
        blocks.append(cur)
for b in blocks:
    if 100 * b["inst"] / tot >= min_share:
        ops = []
        for i in range(b["start"], b["end"] + 1):
            t = data[i][iS].split()
            ops.append((t[1] if t[0].startswith("@") else t[0]).split(".")[0])
        c = collections.Counter(ops).most_common(7)
        print(f"[{b['start']:4d}-{b['end']:4d}] n={b['n']:3d} exec={b['e']:9d} share={100*b['inst']/tot:5.1f}% "
              f"samples={100*b['smp']/totS:5.1f}% thr={b['thr']:>5s} {c}")
