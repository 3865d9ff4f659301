"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).
usage: python scripts/launch_summary.py LIST.csv "header comment" [first_id last_id] > profiles/xxx_summary.csv
first_id / last_id: restrict to a range of launch IDs (e.g. the second frame of scripts/one_frame.py)."""
import csv, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit() or (r and r[0] == "ID")]
hdr = rows[0]
iid, ik, iv, iu = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
lo = int(sys.argv[3]) if len(sys.argv) > 3 else -1
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 30
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
acc = OrderedDict()
for r in rows[1:]:
    if not (lo <= int(r[iid]) <= hi):
        continue
    name = re.sub(r"^(void )?", "", r[ik].split("(")[0])
    name = re.sub(r"^(<unnamed>|.anonymous namespace.)::", "", name)[:120]
    ms = float(r[iv].replace(",", "")) * scale[r[iu]]
    n, t = acc.get(name, (0, 0.0))
    acc[name] = (n + 1, t + ms)
tot = sum(t for _, t in acc.values())
print(f"# {sys.argv[2]}")
print(f"# (cold-cache, serialised launches: compare SHARES, not absolutes); total {tot:.3f} ms, {sum(n for n, _ in acc.values())} launches")
print("kernel,launches,total_ms,ms_per_launch,share")
for name, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{name},{n},{t:.4f},{t / n:.4f},{t / tot:.4f}")
