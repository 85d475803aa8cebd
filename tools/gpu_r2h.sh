#!/bin/bash
set -x
N=${1:-2}
for st in 10 5 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$st bench.py --gpus $N --steps $st --warmup 3 --no-suite --no-cpu > gpurun_out/r2h_bench_n${N}_s$st.json 2> gpurun_out/r2h_bench_n${N}_s$st.err
grep -h "AssertionError" gpurun_out/r2h_bench_n${N}_s$st.err | head -2
head -c 300 gpurun_out/r2h_bench_n${N}_s$st.json
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --no-suite --no-cpu --gather all > gpurun_out/r2h_bench_n${N}_all.json 2> gpurun_out/r2h_bench_n${N}_all.err
grep -h "AssertionError" gpurun_out/r2h_bench_n${N}_all.err | head -2
head -c 300 gpurun_out/r2h_bench_n${N}_all.json
