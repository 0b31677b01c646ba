"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: python tools/ncu_summary.py launches.csv [first_row last_row]"""
import csv, re, sys, collections
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"): v *= 1e3
        if r["Metric Unit"] in ("ms", "msecond"): v *= 1e6
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((int(r["ID"]), name, v, r["Grid Size"], r["Block Size"]))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
rows = rows[lo:hi]
agg = collections.OrderedDict()
for _, n, v, g, b in rows:
    a = agg.setdefault(n, [0, 0.0, g, b]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("launches %d..%d: %d kernels, %.1f us total (serialised, cold-cache ncu replay times)" % (lo, hi, len(rows), tot / 1e3))
print("%-58s %6s %10s %8s %7s  %s" % ("kernel", "count", "total us", "avg us", "share", "grid/block (first)"))
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-58s %6d %10.1f %8.2f %6.1f%%  %s %s" % (n[:58], a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / tot, a[2], a[3]))
