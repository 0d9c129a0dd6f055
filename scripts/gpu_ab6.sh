#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/ab_probe.py main 32 2>&1 | grep -v "Light object\|BVH has" | tee gpurun_out/ab6.log
timeout 300 python scripts/ab_probe.py main_sponza 32 scene=sponza 2>&1 | grep -v "Light object\|BVH has" | tee -a gpurun_out/ab6.log
timeout 600 python scripts/post_probe.py 8 2>gpurun_out/post_probe.err | grep '^{' | tee gpurun_out/post_probe.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:k_atrous|k_spatial_clamp|k_filter|k_bloom|k_shade_gamma|k_gamma|k_fxaa' -c 40 --csv --log-file gpurun_out/launches_post.csv python scripts/post_probe.py 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_post.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:40]),{})[r[mi]]=r[vi]
seen=set()
for (i,k),m in d.items():
    if k in seen: continue
    seen.add(k); print(k, m)
PY
