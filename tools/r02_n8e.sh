#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc $?"; }
run r02e_c4_n8 NRT_DUMMY=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['n_gpus'], round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms; dev', round(d['device_ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'ms', (d.get('parity') or {}).get('matches_oracle'))
            print('   dev ms per rank', [round(r['device_ms_per_step'],3) for r in d.get('ranks',[])])
PY
