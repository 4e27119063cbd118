#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02f_pytest.log
tail -3 gpurun_out/r02f_pytest.log
{
for l in 1 4; do NRT_LANES=$l timeout 300 python tools/frame_breakdown.py config4; done
for l in 1 4; do NRT_PART=0,8 NRT_LANES=$l timeout 300 python tools/frame_breakdown.py config4; done
} > gpurun_out/r02f_ab.log 2>&1
cat gpurun_out/r02f_ab.log
