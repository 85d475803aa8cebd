#!/bin/bash
# quick loop: GPU tests (optional -k expression as $1) and probes given as the remaining arguments
K="$1"; shift
mkdir -p gpurun_out
if [ -n "$K" ]; then timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -k "$K" > gpurun_out/quick_pytest.log 2>&1; else timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 --tb=short > gpurun_out/quick_pytest.log 2>&1; fi
echo "pytest rc=$?"; tail -25 gpurun_out/quick_pytest.log
if [ $# -gt 0 ]; then python tools/probe_hot.py "$@" > gpurun_out/quick_probe.jsonl 2> gpurun_out/quick_probe.err; cat gpurun_out/quick_probe.jsonl; tail -3 gpurun_out/quick_probe.err; fi
