"""Per-launch table from an `ncu --csv` log with several metrics per kernel: name, duration, DRAM read/write MB, L2 hit %, issue %."""
import csv
import sys
from collections import OrderedDict

rows = OrderedDict()
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    key = r["ID"]
    d = rows.setdefault(key, {"name": r["Kernel Name"][:60]})
    v = r["Metric Value"].replace(",", "")
    try:
        v = float(v)
    except ValueError:
        continue
    unit = r["Metric Unit"]
    n = r["Metric Name"]
    if n.startswith("gpu__time"):
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    elif n.startswith("dram__bytes"):
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
        d["rd" if "read" in n else "wr"] = v * scale
    elif n.startswith("lts__t_sector_hit"):
        d["l2hit"] = v
    elif n.startswith("smsp__issue"):
        d["issue"] = v
tot = {"us": 0, "rd": 0, "wr": 0}
for k, d in rows.items():
    print(f'{d["name"]:60s} {d.get("us", 0):9.1f} us  rd {d.get("rd", 0):8.1f} MB  wr {d.get("wr", 0):8.1f} MB  L2hit {d.get("l2hit", 0):5.1f}%  issue {d.get("issue", 0):5.1f}%')
    for t in tot:
        tot[t] += d.get(t, 0)
print(f'{"TOTAL":60s} {tot["us"]:9.1f} us  rd {tot["rd"]:8.1f} MB  wr {tot["wr"]:8.1f} MB')
