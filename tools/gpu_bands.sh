#!/bin/bash
# band-sharded (strong scaling) bench at N = world size of this call
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 280 python bench.py --mode bands --steps 20 --warmup 3 > gpurun_out/r1e_bench_n1_bands.json 2> gpurun_out/r1e_bands_n1.err; echo rc=$?; tail -3 gpurun_out/r1e_bands_n1.err
else
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --mode bands --steps 20 --warmup 3 > gpurun_out/r1e_bench_n${N}_bands.json 2> gpurun_out/r1e_bands_n$N.err; echo rc=$?; tail -3 gpurun_out/r1e_bands_n$N.err
fi
cut -c1-300 gpurun_out/r1e_bench_n${N}_bands.json
