#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tl_*.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
{
for cfg in "1 1 0" "1 1 4096" "2 1 4096" "3 1 4096" "4 2 4096" "2 1 0" "3 1 0"; do
set -- $cfg
echo "=== LANES $1 HEAVY $2 FORK_MIN $3 part 0,8"
NRT_FORK_MIN=$3 NRT_LANES=$1 NRT_HEAVY_LANES=$2 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha" | tail -3
done
echo "=== no feedback 4 lanes fork"
NRT_LANE_FEEDBACK=0 NRT_LANES=4 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha" | tail -3
echo "=== no feedback 2 lanes fork"
NRT_LANE_FEEDBACK=0 NRT_LANES=2 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha" | tail -3
for cfg in "4 2 4096 1" "4 2 4096 0" "1 1 4096 0" "2 1 4096 0" "4 1 0 0" "2 1 4096 1"; do
set -- $cfg
echo "=== FULL LANES $1 HEAVY $2 FORK_MIN $3 FEEDBACK $4"
NRT_LANE_FEEDBACK=$4 NRT_FORK_MIN=$3 NRT_LANES=$1 NRT_HEAVY_LANES=$2 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha" | tail -3
done
NRT_TIMELINE=gpurun_out/tl_part8.txt NRT_LANES=1 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 > /dev/null 2>&1
} > gpurun_out/r02y.log 2>&1
cut -c1-250 gpurun_out/r02y.log
