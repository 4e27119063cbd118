#!/bin/bash
# session 3, call i: e2e with / without staggered lane priorities (host framebuffer)
mkdir -p gpurun_out
for st in 0 1 0 1; do
NRT_STAGGER=$st timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s3i_st$st.json 2> gpurun_out/s3i_st$st.err
python - <<PY
import json
for l in open('gpurun_out/s3i_st$st.json'):
    if l.startswith('{'):
        d=json.loads(l); print('stagger $st:', round(d['ms_per_step'],3), 'ms resident; e2e', round(d['e2e']['ms_per_step'],3), 'ms', d['parity']['matches_oracle'])
PY
done
