#!/bin/bash
# N ranks: the driver's default line (+ sub-lines) and BASELINE config 5 (64 pairs of 1280x720 / 64 disparities per rank)
N=${1:-8}; T=${2:-r2q}
mkdir -p gpurun_out
bash tools/gpu_scale.sh $N $T
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $N --workload 720p_d64_w9 --pairs 64 --steps 10 --warmup 3 --no-suite > gpurun_out/${T}_bench_n${N}_720p_x64_config5.json 2> gpurun_out/${T}_bench_n${N}_720p.err; echo "config5 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_n${N}_720p_x64_config5.json') if l.startswith('{')][-1])
print('config5', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_check']['ok'])
PY
