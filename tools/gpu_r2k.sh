#!/bin/bash
set -x
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2k_topo.txt 2>&1; nproc >> gpurun_out/r2k_topo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2k_bench_n${N}.json 2> gpurun_out/r2k_bench_n${N}.err
grep -h "Error" gpurun_out/r2k_bench_n${N}.err | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 --workload 720p_d64_w9 --pairs 64 --no-suite > gpurun_out/r2k_720p_x64_n${N}.json 2> gpurun_out/r2k_720p_x64_n${N}.err
grep -h "Error" gpurun_out/r2k_720p_x64_n${N}.err | head -5
python - <<PY
import json
for f in ('gpurun_out/r2k_bench_n${N}.json','gpurun_out/r2k_720p_x64_n${N}.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['value'], d['ms_per_step'], d['e2e'], d['parity_check'])
        print(json.dumps(d.get('gather_variants'))[:900])
        b=d.get('bands')
        if b: print('bands', b['value'], b['ms_per_step'], b['config']['launch'], b['parity_check'])
    except Exception as e: print(f,'ERR',e)
PY
