#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
{
NRT_LANES=1 timeout 600 python tools/frame_breakdown.py config5s
NRT_LANES=4 NRT_PART=0,8 timeout 600 python tools/frame_breakdown.py config5
NRT_LANES=4 timeout 300 python tools/frame_breakdown.py config4 config1
} > gpurun_out/r02q.log 2>&1
grep -v "fb sha" gpurun_out/r02q.log
