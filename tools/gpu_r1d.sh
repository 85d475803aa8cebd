#!/bin/bash
# Round-1d GPU session: parity tests, the bench lines, the host-pipeline sweep.  Run under gpurun from the repo root.
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r1d_bench_n1_4k_ssd.json 2> gpurun_out/r1d_bench_n1_4k_ssd.err; echo "bench rc=$?"; cat gpurun_out/r1d_bench_n1_4k_ssd.json; tail -5 gpurun_out/r1d_bench_n1_4k_ssd.err
timeout 200 python tools/e2e_sweep.py --workload 4k_d256_w11 --pairs 1 2 4 --bands 0 1 2 4 8 > gpurun_out/r1d_e2e_sweep_4k.jsonl 2>&1; cat gpurun_out/r1d_e2e_sweep_4k.jsonl
timeout 200 python tools/e2e_sweep.py --workload 720p_d64_w9 --pairs 4 16 64 --bands 0 1 > gpurun_out/r1d_e2e_sweep_720p.jsonl 2>&1; cat gpurun_out/r1d_e2e_sweep_720p.jsonl
timeout 300 python bench.py --steps 20 --warmup 3 --cost ncc > gpurun_out/r1d_bench_n1_4k_ncc.json 2> gpurun_out/r1d_bench_n1_4k_ncc.err; cat gpurun_out/r1d_bench_n1_4k_ncc.json
timeout 300 python bench.py --steps 20 --warmup 3 --workload 720p_d64_w9 --pairs 16 > gpurun_out/r1d_bench_n1_720p_x16_ssd.json 2> gpurun_out/r1d_bench_n1_720p.err; cat gpurun_out/r1d_bench_n1_720p_x16_ssd.json
timeout 300 python bench.py --steps 20 --warmup 3 --workload 1080p_d128_w9 > gpurun_out/r1d_bench_n1_1080p_ssd.json 2> gpurun_out/r1d_bench_n1_1080p.err; cat gpurun_out/r1d_bench_n1_1080p_ssd.json
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1d_bench_reference_arm.json 2>&1; cat gpurun_out/r1d_bench_reference_arm.json
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 --cost ncc > gpurun_out/r1d_bench_reference_arm_ncc.json 2>&1; cat gpurun_out/r1d_bench_reference_arm_ncc.json
