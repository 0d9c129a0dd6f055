"""Rank CUDA source lines of one kernel by warp-stall samples.
usage: python scripts/ncu_lines.py report.ncu-rep kernel_regex [launch_skip] [top]"""
import csv, subprocess, sys, io, os
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name",) or hdr is None: continue
    if r[0] == "": continue
    d = dict(zip(hdr[2:], r[2:]))  # metrics follow (dup 'Source' col for sass is '-')
    try: samples = int(r[4])
    except Exception: continue
    stalls = {k: int(v) for k, v in zip(hdr, r) if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    inst = r[7]; thr = r[10]
    lines.append((samples, cur_file, r[0], r[1].strip()[:90], inst, thr, stalls))
tot = sum(l[0] for l in lines)
print("total samples", tot)
agg = {}
for l in lines:
    for k, v in l[6].items(): agg[k] = agg.get(k, 0) + v
print("stall mix:", ", ".join("%s %.1f%%" % (k, 100 * v / max(1, sum(agg.values()))) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for l in sorted(lines, key=lambda l: -l[0])[:top]:
    st = ",".join("%s:%d" % (k.replace("stall_", ""), v) for k, v in sorted(l[6].items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% %-20s:%-4s inst=%-9s thr=%-4s %s | %s" % (100 * l[0] / max(tot, 1), l[1], l[2], l[4], l[5], l[3], st))
