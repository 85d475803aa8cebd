#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu > gpurun_out/b3_4k.json 2> gpurun_out/b3_4k.err; echo "rc=$?"
timeout 300 python bench.py --no-cpu --workload 1080p_d128_w9 > gpurun_out/b3_1080p.json 2> gpurun_out/b3_1080p.err; echo "rc=$?"
timeout 300 python bench.py --no-cpu --workload 720p_d64_w9 --pairs 16 > gpurun_out/b3_720p.json 2> gpurun_out/b3_720p.err; echo "rc=$?"
timeout 300 python bench.py --mode bands > gpurun_out/b3_bands.json 2> gpurun_out/b3_bands.err; echo "rc=$?"
