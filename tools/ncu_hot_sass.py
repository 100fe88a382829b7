"""Top SASS instructions by warp-stall samples from an .ncu-rep captured with --import-source on."""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
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
        data.append((float(r[ci]), n, r[si].strip(), r[ei]))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
print(f"{rep}: {len(data)} instructions, {tot:.0f} stall samples")
for v, n, src, ex in sorted(data, reverse=True)[:top]:
    print(f"{v:8.0f} {100 * v / tot:5.1f}%  #{n:5d} exec {ex:>9s}  {src[:100]}")
