#!/bin/bash
# round-2 experiment 1: GPU tests, then config-4 frame time per path mode / occupancy build
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
{
for p in 0 1 2; do NRT_PATH=$p timeout 300 python tools/frame_breakdown.py config4 config3 config2; done
for v in 3 4; do NRT_LIB=$PWD/nim_raytracer_b200/csrc/libnrt_occ$v.so NRT_PATH=1 timeout 300 python tools/frame_breakdown.py config4; done
for part in 0,8 3,8; do NRT_PART=$part NRT_PATH=1 timeout 300 python tools/frame_breakdown.py config4; NRT_PART=$part NRT_PATH=0 timeout 300 python tools/frame_breakdown.py config4; NRT_PART=$part NRT_PATH=2 timeout 300 python tools/frame_breakdown.py config4; done
} > gpurun_out/r02a_ab.log 2>&1
cat gpurun_out/r02a_ab.log
