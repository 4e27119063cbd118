#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
{
NRT_LANES=1 timeout 300 python tools/frame_breakdown.py config4
NRT_LANES=4 timeout 300 python tools/frame_breakdown.py config4 config3 config2
NRT_LANES=4 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4
} > gpurun_out/r02t.log 2>&1
grep -v "fb sha" gpurun_out/r02t.log
