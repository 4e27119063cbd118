#!/bin/bash
mkdir -p gpurun_out
{
for tb in 32768 131072 400000; do NRT_TAIL_BELOW=$tb NRT_LANES=4 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4; done
for l in 2 3 6; do NRT_LANES=$l NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4; done
NRT_TAIL_BELOW=131072 NRT_LANES=4 timeout 300 python tools/frame_breakdown.py config4
NRT_TAIL_BELOW=1000000 NRT_LANES=4 timeout 300 python tools/frame_breakdown.py config4
} > gpurun_out/r02m_ab.log 2>&1
grep -v "fb sha" gpurun_out/r02m_ab.log | grep -v "^   "
