#!/bin/bash
# 2 GPUs: multi-rank bench (pairs, gather root + variants + bands subline), bands mode standalone, sharding tests
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2g_bench_n${N}.json 2> gpurun_out/r2g_bench_n${N}.err
tail -3 gpurun_out/r2g_bench_n${N}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --mode bands > gpurun_out/r2g_bands_n${N}.json 2> gpurun_out/r2g_bands_n${N}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --mode bands --no-graph > gpurun_out/r2g_bands_nograph_n${N}.json 2> gpurun_out/r2g_bands_nograph_n${N}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --workload 720p_d64_w9 --pairs 64 --no-suite > gpurun_out/r2g_720p_x64_n${N}.json 2> gpurun_out/r2g_720p_x64_n${N}.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2g_*_n${N}.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('parity_check'), d['config'].get('launch'))
        for k in ('gather_variants','bands'):
            if k in d: print('   ',k, json.dumps(d[k])[:600])
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
