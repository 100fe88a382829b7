"""Extract per-launch DRAM traffic of named kernels from an `ncu --csv` log (dram__bytes_read.sum / dram__bytes_write.sum) into a
small JSON that bench.py reads for `roofline.traffic` (keyed by kernel + shape, so the figure can never silently go stale: when the
kernel or the shape changes the key is missing and bench.py reports null).
Usage: python tools/ncu_traffic.py <ncu.csv> <out.json> "<key>::<kernel substring>@<launch index among matches>" ...
"""
import csv
import json
import sys


def main():
    path, out = sys.argv[1], sys.argv[2]
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per_id = {}
    for r in csv.DictReader(lines):
        d = per_id.setdefault(r["ID"], {"name": r["Kernel Name"]})
        n, unit = r["Metric Name"], r["Metric Unit"]
        if n.startswith("dram__bytes"):
            v = float(r["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d["read" if "read" in n else "write"] = v
        elif n.startswith("gpu__time"):
            d["us"] = float(r["Metric Value"].replace(",", "")) / (1e3 if unit.startswith("n") else 1)
    launches = [per_id[k] for k in sorted(per_id, key=int)]
    res = json.load(open(out)) if len(sys.argv) > 3 and __import__("os").path.isfile(out) else {}
    for spec in sys.argv[3:]:
        key, sel = spec.split("::", 1)
        sub, idx = sel.rsplit("@", 1)
        m = [l for l in launches if sub in l["name"]]
        l = m[int(idx)]
        res[key] = {"dram_read_bytes": int(l.get("read", 0)), "dram_write_bytes": int(l.get("write", 0)),
                    "traffic_bytes": int(l.get("read", 0) + l.get("write", 0)), "ncu_us": l.get("us"), "source": path}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
