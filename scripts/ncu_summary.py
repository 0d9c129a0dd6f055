"""Key metrics per captured launch of an .ncu-rep (raw page)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_bytes.sum", "l2_bytes"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%")]
idx = {h: i for i, h in enumerate(hdr)}
for d in data:
    name = d[idx["Kernel Name"]][:60]
    parts = []
    for k, lab in want:
        if k in idx:
            v = d[idx[k]]
            try: v = "%.4g" % float(v.replace(",", ""))
            except Exception: pass
            parts.append("%s=%s%s" % (lab, v, units[idx[k]] if lab in ("time", "dram_rd", "dram_wr", "l2_bytes") else ""))
    print(d[idx["ID"]], name, "|", " ".join(parts))
