#!/bin/bash
# Round-2 GPU pass B: parity incl. the float running-sum kernels, headline bench with suite, launch lists of the small workloads.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_4k.json 2> gpurun_out/r2b_bench_4k.err
for wl in ps2_pair1_511x640_d96_w15; do
  for cost in ssd ncc; do
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2b_launches_${wl}_${cost}.csv \
      python bench.py --workload $wl --cost $cost --pairs 1 --steps 3 --warmup 3 --no-cpu --no-suite --no-parity > gpurun_out/r2b_ncu_${wl}_${cost}.log 2>&1
  done
done
head -c 400 gpurun_out/r2b_bench_4k.json
