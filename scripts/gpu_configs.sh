#!/bin/bash
# the other BASELINE configs at full size: stage timings (parity is covered by the tests)
mkdir -p gpurun_out
( echo "== config 1 cornell 512x512 64spp"; timeout 300 python scripts/perf_probe.py cornell 0 512 512 64
  echo "== config 2 sponza-scale 260K + sky 1080p 256spp"; timeout 600 python scripts/perf_probe.py sponza 260000 1920 1080 256
  echo "== config 4 texture-heavy 500K tris, 32 x 2048^2 albedo+normal maps, 4K, 64 of its 4096 spp (one GPU's share of 8 would be 512)"; timeout 900 python scripts/perf_probe.py texture 500000 3840 2160 64
  echo "== config 5 5M tris 4K primary"; timeout 900 python scripts/perf_probe.py five 5000000 3840 2160 0 count
) > gpurun_out/configs.log 2>&1
cat gpurun_out/configs.log | grep -v "Light object"
