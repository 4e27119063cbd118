#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc $?"
}
run r02b_c4_n8_plan NRT_DUMMY=1
run r02b_c4_n8_rr4 NRT_LANE_FEEDBACK=0 NRT_LANES=4
run r02b_c4_n8_plan2 NRT_LANES=2
env timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29590 bench.py --gpus 8 --workload config5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_c5_n8.json 2> gpurun_out/r02b_c5_n8.err; echo "c5 rc $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['n_gpus'], round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'ms', (d.get('parity') or {}).get('matches_oracle'))
PY
tail -3 gpurun_out/r02b_*.err
