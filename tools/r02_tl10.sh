#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
{
echo "=== FULL default"
timeout 300 python tools/frame_breakdown.py config4 config3 config2 config1 2>&1 | grep -v "fb sha"
echo "=== part 0,8"
NRT_TAIL_FILL=4 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha"
echo "=== config5 part 0,8"
NRT_PART=0,8 timeout 600 python tools/frame_breakdown.py config5 2>&1
} > gpurun_out/r02zd.log 2>&1
cut -c1-330 gpurun_out/r02zd.log
