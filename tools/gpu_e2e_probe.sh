#!/bin/bash
# e2e diagnostics on one GPU: the CV_32FC1 / u8 batch entry points over host threads,
# then one traced call per input kind (STEREO_PIPE_TRACE: per work item and
# band, when its upload / compute / download finished).
#   bash tools/gpu_e2e_probe.sh <tag> [workload] [pairs]
mkdir -p gpurun_out
T=${1:-e2e}; W=${2:-4k_d256_w11}; P=${3:-4}
: > gpurun_out/${T}_sweep.jsonl
python tools/e2e_sweep.py --workload $W --pairs $P --bands 0 --threads 8 16 --steps 10 --kinds f32 >> gpurun_out/${T}_sweep.jsonl 2>> gpurun_out/${T}_sweep.err
python tools/e2e_sweep.py --workload $W --pairs $P 1 --bands 0 --steps 10 --kinds u8 >> gpurun_out/${T}_sweep.jsonl 2>> gpurun_out/${T}_sweep.err
python tools/e2e_sweep.py --workload $W --pairs 1 --bands 0 --steps 10 --kinds f32 >> gpurun_out/${T}_sweep.jsonl 2>> gpurun_out/${T}_sweep.err
cut -c1-400 gpurun_out/${T}_sweep.jsonl | sed 's/"workload[^k]*"kind"/"kind"/'
STEREO_PIPE_TRACE=1 python tools/e2e_sweep.py --workload $W --pairs $P --bands 0 --steps 1 > /dev/null 2> gpurun_out/${T}_trace.txt
awk -v p="$P pairs" '/\[pipe\] /{if (index($0, p)) n++} n==3||n==6' gpurun_out/${T}_trace.txt | head -60
