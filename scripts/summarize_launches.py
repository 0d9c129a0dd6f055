"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    rows.append((int(r["ID"]), name, ns, r["Grid Size"], r["Block Size"]))
tot = sum(r[2] for r in rows)
agg = collections.OrderedDict()
for _, n, ns, g, b in rows:
    a = agg.setdefault(n, [0, 0.0, g, b]); a[0] += 1; a[1] += ns
print("%-32s %8s %12s %8s %10s  %s" % ("kernel", "launches", "total_ms", "share", "avg_us", "grid x block"))
for n, (c, ns, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-32s %8d %12.3f %7.1f%% %10.1f  %s x %s" % (n, c, ns / 1e6, 100 * ns / tot, ns / c / 1e3, g, b))
print("%-32s %8d %12.3f" % ("TOTAL", len(rows), tot / 1e6))
if "--list" in sys.argv:
    for i, n, ns, g, b in rows: print(i, n, "%.1f us" % (ns / 1e3))
