#!/bin/bash
# the driver's scaling run at N ranks: default line (pairs, gather to root, sub-lines) + reference arm
N=${1:-2}; T=${2:-r2q}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${T}_ref_n$N.json 2> gpurun_out/${T}_ref_n$N.err; echo "ref rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_n$N.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_check']['ok'], {k:v['value'] for k,v in d.get('gather_variants',{}).items()}, (d.get('bands') or {}).get('value'), (d.get('bands') or {}).get('ms_per_step'))
PY
