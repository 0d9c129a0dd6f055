"""Rank CUDA source lines of one kernel by warp-level instructions executed."""
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
    if hdr is None or r[0] in ("Function Name", ""): continue
    try: inst, thr = int(r[7]), int(r[8])
    except Exception: continue
    lines.append((inst, thr, cur_file, r[0], r[1].strip()[:100]))
tot = sum(l[0] for l in lines); tthr = sum(l[1] for l in lines)
print("warp instructions %d, thread instructions %d, avg threads %.1f" % (tot, tthr, tthr / max(tot, 1)))
for l in sorted(lines, key=lambda l: -l[0])[:top]:
    print("%5.1f%% inst=%-10d thr/inst=%4.1f %-18s:%-4s %s" % (100 * l[0] / tot, l[0], l[1] / max(l[0], 1), l[2], l[3], l[4]))
