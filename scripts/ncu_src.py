"""Per source line: warp instructions, threads/inst and stall samples of one kernel launch (ncu source page).
usage: python scripts/ncu_src.py report.ncu-rep demangled_kernel_regex [launch_skip] [top] [sort=inst|samples]"""
import csv, subprocess, sys, io, os, collections
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
key = sys.argv[5] if len(sys.argv) > 5 else "inst"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name-base", "demangled",
                      "--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr = None, None
agg = collections.OrderedDict()
stall_tot = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("Function Name", ""): continue
    d = {}
    for k, v in zip(hdr, r):
        if k not in d: d[k] = v
    try:
        inst = int(d["Instructions Executed"]); thr = int(d["Thread Instructions Executed"]); smp = int(d["# Samples"])
    except Exception: continue
    a = agg.setdefault((cur_file, r[0]), [0, 0, 0, r[1].strip()[:100], collections.Counter(), 0])
    a[0] += inst; a[1] += thr; a[2] += smp; a[5] += 1
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v):
            a[4][k[6:]] += int(v); stall_tot[k[6:]] += int(v)
tot = sum(a[0] for a in agg.values()); tthr = sum(a[1] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print("warp inst %d  thread inst %d  avg threads %.2f  samples %d" % (tot, tthr, tthr / max(tot, 1), ts))
print("stall mix:", ", ".join("%s %.1f%%" % (k, 100 * v / max(1, sum(stall_tot.values()))) for k, v in stall_tot.most_common(9)))
idx = 0 if key == "inst" else 2
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][idx])[:top]:
    st = ",".join("%s:%d" % kv for kv in a[4].most_common(3))
    print("%5.1f%%i %5.1f%%s sass=%-3d thr/inst=%4.1f %-18s:%-4s %s | %s" % (100 * a[0] / max(tot, 1), 100 * a[2] / max(ts, 1), a[5], a[1] / max(a[0], 1), f, ln, a[3], st))
