#!/bin/bash
# the other BASELINE configs at full size: stage timings (parity is covered by the tests at small sizes)
mkdir -p gpurun_out
( echo "== config 1 cornell 512x512 64spp"; timeout 300 python scripts/perf_probe.py cornell 0 512 512 64
  echo "== config 2 sponza-scale 260K + sky 1080p 64spp"; timeout 600 python scripts/perf_probe.py sponza 260000 1920 1080 64
  echo "== config 5 5M tris 4K primary"; timeout 900 python scripts/perf_probe.py five 5000000 3840 2160 0 count
) > gpurun_out/configs.log 2>&1
cat gpurun_out/configs.log | grep -v "Light object"
