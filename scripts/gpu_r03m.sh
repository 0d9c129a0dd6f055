#!/bin/bash
# round 2, call 3m: the sweep-SAH builder at the ends of its range: scenes of 1, 2, 3, 5 triangles (test) and the 5 M-triangle scene
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tree.py -m gpu -x -q -k "handful or first_needed" ) 2>&1 | tail -3
timeout 600 python scripts/big_tree_check.py 2>/dev/null | tee gpurun_out/r03m_sweep_sah_5m_triangles.json
