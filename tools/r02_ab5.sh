#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02e_pytest.log
tail -5 gpurun_out/r02e_pytest.log
{
for l in 1 2 4 8; do NRT_LANES=$l timeout 300 python tools/frame_breakdown.py config4; done
for l in 1 2 4 8; do NRT_PART=0,8 NRT_LANES=$l timeout 300 python tools/frame_breakdown.py config4; done
for l in 1 4; do NRT_PART=3,8 NRT_LANES=$l timeout 300 python tools/frame_breakdown.py config4; done
NRT_LANES=4 timeout 300 python tools/frame_breakdown.py config3 config2
} > gpurun_out/r02e_ab.log 2>&1
grep -v "^   [A-Za-z]" gpurun_out/r02e_ab.log
