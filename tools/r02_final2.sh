#!/bin/bash
# end-of-round-2 measurement set on one B200 (after the by-value scene header and the 16-flags select):
# ncu --set full of one frame's kernels -> kernel_traffic.json, bench line, reference arm, ncu launch list of the same bench command
mkdir -p gpurun_out
NRT_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_for_each_stats|k_mesh_prefilter|k_prefilter_bounds|k_gate_write|k_produce_gate|k_gate_flags|k_finalize|k_path_warp|k_for_each_counted|k_sel_" --launch-count 32 -o gpurun_out/r02b_ncu_frame_config4 -f python tools/frame_breakdown.py config4 > gpurun_out/r02b_ncu_full.log 2>&1; echo "ncu full rc $?"
python tools/make_kernel_traffic.py gpurun_out/r02b_ncu_frame_config4.ncu-rep config4 132710400 profiles/kernel_traffic.json > gpurun_out/r02b_kernel_traffic.log 2>&1; echo "traffic rc $?"
cp profiles/kernel_traffic.json gpurun_out/kernel_traffic.json
python tools/ncu_metrics.py gpurun_out/r02b_ncu_frame_config4.ncu-rep 12 > gpurun_out/r02b_ncu_frame_config4.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_config4_n1.json 2> gpurun_out/r02b_bench_config4_n1.err; echo "bench rc $?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_reference_arm.json 2>/dev/null; echo "ref rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02b_launches_bench_config4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_ncu_launches.log 2>&1; echo "ncu launches rc $?"
for w in config2 config3; do timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_${w}_n1.json 2>/dev/null; done
rm -f gpurun_out/r02b_ncu_frame_config4.ncu-rep.tmp
ls -la gpurun_out | tail -14
