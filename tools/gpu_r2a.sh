#!/bin/bash
# Round-2 GPU pass A: parity (whole -m gpu suite), then bench lines for the headline and the ps2-shaped workloads.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_4k.json 2> gpurun_out/r2a_bench_4k.err
for wl in ps2_pair1_511x640_d96_w15 ps2_pair0_128_d4_w13 ps2_pair2_529x640_d81_w15; do
  for cost in ssd ncc; do
    for pairs in 1 4; do
      timeout 300 python bench.py --workload $wl --cost $cost --pairs $pairs --steps 50 --warmup 5 --no-cpu > gpurun_out/r2a_bench_${wl}_${cost}_x${pairs}.json 2> gpurun_out/r2a_bench_${wl}_${cost}_x${pairs}.err
    done
  done
done
STEREO_FAST_NW=8 timeout 300 python bench.py --workload ps2_pair1_511x640_d96_w15 --cost ssd --pairs 1 --steps 50 --warmup 5 --no-cpu > gpurun_out/r2a_bench_pair1_ssd_x1_nw8.json 2>&1
STEREO_FUSE_PAIRS=0 timeout 300 python bench.py --workload ps2_pair1_511x640_d96_w15 --cost ssd --pairs 1 --steps 50 --warmup 5 --no-cpu > gpurun_out/r2a_bench_pair1_ssd_x1_unfused.json 2>&1
timeout 300 python bench.py --workload 720p_d64_w9 --pairs 16 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_bench_720p.json 2> gpurun_out/r2a_bench_720p.err
timeout 300 python bench.py --workload 1080p_d128_w9 --pairs 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_bench_1080p.json 2> gpurun_out/r2a_bench_1080p.err
head -c 600 gpurun_out/r2a_bench_4k.json
