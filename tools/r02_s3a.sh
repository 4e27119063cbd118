#!/bin/bash
# session 3, call a: HEAD check — GPU tests, the default bench line, full frame + 1/8 frame breakdowns
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/s3a_bench_config4_n1.json 2> gpurun_out/s3a_bench_config4_n1.err; echo "bench rc $?"
{
echo "=== FULL default"
timeout 300 python tools/frame_breakdown.py config4 config3 config2 2>&1
for k in 0 3 7; do
echo "=== part $k,8"
NRT_TRACE_LANES=1 NRT_PART=$k,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "^\[lanes\]" 
done
echo "=== part 0,8 one lane (kernel times are per family)"
NRT_LANES=1 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1
} > gpurun_out/s3a.log 2>&1
cut -c1-400 gpurun_out/s3a.log
python - <<'PY'
import json
for l in open('gpurun_out/s3a_bench_config4_n1.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity'))
PY
tail -3 gpurun_out/s3a_bench_config4_n1.err
