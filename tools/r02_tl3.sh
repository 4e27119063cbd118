#!/bin/bash
mkdir -p gpurun_out
{
for tb in 32768 100000 300000 3000000; do
for ln in 1 2; do
echo "=== TAIL_BELOW $tb LANES $ln part 0,8"
NRT_TAIL_BELOW=$tb NRT_LANES=$ln NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha"
done
done
for tb in 32768 100000 600000; do
echo "=== TAIL_BELOW $tb full"
NRT_TAIL_BELOW=$tb timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha"
done
} > gpurun_out/r02w.log 2>&1
cut -c1-330 gpurun_out/r02w.log
