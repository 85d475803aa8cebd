#!/bin/bash
# e2e check at N ranks: the default line without suite / cpu legs
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --no-suite --no-cpu --no-parity > gpurun_out/e2e_n1.json 2> gpurun_out/e2e_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 --no-suite --no-cpu --no-parity > gpurun_out/e2e_n$N.json 2> gpurun_out/e2e_n$N.err
fi
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/e2e_n$N.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], json.dumps(d['e2e']))
PY
