"""Per-kernel SASS evidence: counts of the Blackwell-native mnemonics (tcgen05 -> UTC*MMA / LDTM, TMA -> UTMALDG / UTMASTG /
UBLKCP, mbarrier -> SYNCS) in liblws_b200.so.  Usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "lwsnet_b200/lib/liblws_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS", "SHFL"]
cur, counts, total = None, collections.OrderedDict(), {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if cur and m:
        op = m.group(1).split(".")[0]
        total[cur] += 1
        for k in keys:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# {lib}: SASS mnemonic counts per kernel (sm_100a)")
print("kernel".ljust(64) + "instr " + " ".join(k.rjust(7) for k in keys))
for k, c in counts.items():
    print(k[:63].ljust(64) + f"{total[k]:5d} " + " ".join(str(c.get(x, 0)).rjust(7) for x in keys))
