#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
{
for p in 1 2; do NRT_PATH=$p timeout 300 python tools/frame_breakdown.py config4 config3 config2; done
NRT_LIB=$PWD/nim_raytracer_b200/csrc/libnrt_t2.so NRT_PATH=1 timeout 300 python tools/frame_breakdown.py config4
NRT_LIB=$PWD/nim_raytracer_b200/csrc/libnrt_t2.so NRT_PATH=2 timeout 300 python tools/frame_breakdown.py config4
for part in 0,8 3,8; do for p in 1 2; do NRT_PART=$part NRT_PATH=$p timeout 300 python tools/frame_breakdown.py config4; done; done
} > gpurun_out/r02b_ab.log 2>&1
cat gpurun_out/r02b_ab.log
