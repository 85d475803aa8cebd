#!/bin/bash
# two quick bench lines: unfused SSD and NCC (kernel-variant experiments)
mkdir -p gpurun_out
STEREO_FUSE_PAIRS=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --cost ncc > gpurun_out/bench_ncc.json 2> gpurun_out/bench_ncc.err; echo "rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --workload 1080p_d128_w9 > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err; echo "rc=$?"
tail -3 gpurun_out/bench_unfused.err gpurun_out/bench_ncc.err
