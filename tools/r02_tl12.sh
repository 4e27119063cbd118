#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
{
echo "=== FULL default"
timeout 300 python tools/frame_breakdown.py config4 config3 config2 2>&1 | grep -v "fb sha"
echo "=== FULL no tables"
NRT_GEN_TABLES=0 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha"
echo "=== FULL FB at 5 CTAs/SM"
NRT_LIB=/root/repo/tools/ab/libnrt_fb5.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha"
echo "=== part 0,8"
NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha"
} > gpurun_out/r02zf.log 2>&1
cut -c1-330 gpurun_out/r02zf.log
