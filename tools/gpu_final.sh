#!/bin/bash
# Round-end GPU session at N=1: parity tests, smoke, every bench line, ncu launch list + full capture.
T=${1:-r1f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 300 python bench.py > gpurun_out/${T}_bench_n1_4k_ssd.json 2> gpurun_out/${T}_err1.log; echo "bench rc=$?"
timeout 300 python bench.py --cost ncc > gpurun_out/${T}_bench_n1_4k_ncc.json 2> gpurun_out/${T}_err2.log; echo "bench ncc rc=$?"
timeout 300 python bench.py --workload 1080p_d128_w9 > gpurun_out/${T}_bench_n1_1080p_ssd.json 2> gpurun_out/${T}_err3.log; echo "bench 1080p rc=$?"
timeout 300 python bench.py --workload 1080p_d128_w9 --cost ncc > gpurun_out/${T}_bench_n1_1080p_ncc.json 2> gpurun_out/${T}_err4.log; echo "bench 1080p ncc rc=$?"
timeout 300 python bench.py --workload 720p_d64_w9 --pairs 16 > gpurun_out/${T}_bench_n1_720p_x16_ssd.json 2> gpurun_out/${T}_err5.log; echo "bench 720p rc=$?"
timeout 300 python bench.py --mode bands > gpurun_out/${T}_bench_n1_bands.json 2> gpurun_out/${T}_err6.log; echo "bench bands rc=$?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2>&1
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 --cost ncc > gpurun_out/${T}_bench_reference_arm_ncc.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${T}_launches_bench.log 2>&1; echo "ncu bench launches rc=$?"
bash tools/gpu_ncu.sh "" ${T}_fused
bash tools/gpu_ncu.sh "--cost ncc" ${T}_ncc
tail -2 gpurun_out/${T}_err*.log
