"""Group the SASS instructions of an .ncu-rep (captured with --import-source on) by execution count -- in a warp-specialised kernel
every role (producer / MMA issuer / epilogue / front end) executes its loop body a characteristic number of times -- and print, per
class, its share of the executed warp instructions and of the warp-stall samples, plus the hottest instructions of each class."""
import csv
import subprocess
import sys
from collections import Counter, defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
si, ci, ei = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) != len(hdr):
        continue
    try:
        data.append((n, r[si].strip(), float(r[ci]), int(r[ei])))
    except ValueError:
        pass
tot_st = sum(d[2] for d in data) or 1
tot_ex = sum(d[3] for d in data) or 1
cls = defaultdict(list)
for d in data:
    cls[d[3]].append(d)
print(f"{rep}: {len(data)} instructions, {tot_ex} warp instructions executed, {tot_st:.0f} stall samples")
for ex, items in sorted(cls.items(), key=lambda kv: -sum(i[2] for i in kv[1]))[:10]:
    st = sum(i[2] for i in items)
    ops = Counter((i[1].split()[1] if i[1].startswith("@") else i[1].split()[0]).split(".")[0] for i in items)
    print(f"-- exec {ex:9d}: {len(items):4d} instr, {100 * ex * len(items) / tot_ex:5.1f}% of executed, {100 * st / tot_st:5.1f}% of stalls; ops {ops.most_common(6)}")
    for n, src, s, e in sorted(items, key=lambda i: -i[2])[:top]:
        print(f"      {s:7.0f} {100 * s / tot_st:5.1f}%  #{n:5d}  {src[:96]}")
