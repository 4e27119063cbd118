#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
{
for gd in 1000000000 8192 4096 2048 1024 512; do NRT_GRID_DIV=$gd NRT_LANES=4 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4; done
for gd in 4096 2048 1024; do NRT_GRID_DIV=$gd NRT_LANES=4 timeout 300 python tools/frame_breakdown.py config4; done
for gd in 1000000000 2048; do NRT_GRID_DIV=$gd NRT_LANES=8 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4; done
} > gpurun_out/r02n_ab.log 2>&1
grep -v "fb sha" gpurun_out/r02n_ab.log | grep -v "^   "
