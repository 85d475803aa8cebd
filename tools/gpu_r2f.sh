#!/bin/bash
set -x
mkdir -p gpurun_out
nproc > gpurun_out/r2f_nproc.txt; lscpu | head -25 >> gpurun_out/r2f_nproc.txt
echo skip pytest

timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_4k.json 2> gpurun_out/r2f_bench_4k.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2f_bench_4k.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'])
for l in d['suite']['ps2']['problems']: print(l['problem'], l['path'], l['ms_per_pair_call'], l['hot_kernel_ms'], l['vs_baseline'])
print(d['suite']['ps2']['all_problems_ms'])
PY
