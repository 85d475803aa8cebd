#!/bin/bash
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -15 gpurun_out/r2i_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2i_bench_n${N}.json 2> gpurun_out/r2i_bench_n${N}.err
grep -h "Error" gpurun_out/r2i_bench_n${N}.err | head -5
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2i_bench_n${N}.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_check'])
print(json.dumps(d.get('gather_variants'))[:900])
print(json.dumps(d.get('bands'))[:1500])
PY
