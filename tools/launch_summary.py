"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, agg, tot = None, collections.OrderedDict(), 0.0
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1000.0 if d["Metric Unit"] == "ns" else (v * 1000.0 if d["Metric Unit"] == "ms" else v)
        key = re.sub(r"\(.*", "", d["Kernel Name"])[:80]
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"# {path}: {tot:.1f} us total")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  x{n:3d}  {k}")
