#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tl_*.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "short_wavefront or lane_plan or path_modes or config4_full_size or golden_config3" 2>&1 | tail -5
{
for cfg in "1 1" "2 1" "3 1" "4 1" "4 2" "3 2"; do
set -- $cfg
echo "=== LANES $1 HEAVY $2 part 0,8"
NRT_TRACE_LANES=1 NRT_LANES=$1 NRT_HEAVY_LANES=$2 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha" | tail -4
done
echo "=== no feedback 4 lanes"
NRT_LANE_FEEDBACK=0 NRT_LANES=4 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha" | tail -3
echo "=== hard tail off, 2 1"
NRT_HARD_TAIL_BELOW=0 NRT_LANES=2 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha" | tail -3
for cfg in "4 2" "4 1" "2 1" "3 1"; do
set -- $cfg
echo "=== FULL LANES $1 HEAVY $2"
NRT_TRACE_LANES=1 NRT_LANES=$1 NRT_HEAVY_LANES=$2 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha" | tail -4
done
NRT_TIMELINE=gpurun_out/tl_part8.txt NRT_LANES=2 NRT_HEAVY_LANES=1 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 > /dev/null 2>&1
} > gpurun_out/r02x.log 2>&1
cut -c1-300 gpurun_out/r02x.log
