#!/bin/bash
# quick GPU check: the tests named in $1 (a -k expression) or all GPU tests, then an optional short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${1:+-k "$1"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
if [ -n "$2" ]; then timeout 300 python bench.py --steps 20 --warmup 3 $2 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cat gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err; fi
