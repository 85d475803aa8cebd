#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
: > gpurun_out/r2d_probe.jsonl
for tag in "" k12w12 k16w12 k16w8 lea5 lea15; do
  echo "{\"variant\": \"$tag\"}" >> gpurun_out/r2d_probe.jsonl
  STEREO_LIB_TAG=$tag python tools/probe_hot.py 2160,3840,256,5,ssd,4 2160,3840,256,5,ssd,2 >> gpurun_out/r2d_probe.jsonl 2>> gpurun_out/r2d_probe.err
done
python tools/probe_hot.py 511,640,96,7,ssd,1 2044,640,96,7,ssd,1 4088,640,96,7,ssd,1 511,640,96,7,ncc,1 511,640,96,7,ssd,1,noisy 2044,640,96,7,ssd,1,noisy 511,640,96,7,ncc,1,noisy 2044,640,96,7,ncc,1,noisy 2160,3840,256,7,ssd,2 2160,3840,256,5,ssd,1,noisy 2160,3840,256,5,ncc,2 >> gpurun_out/r2d_probe.jsonl 2>> gpurun_out/r2d_probe.err
cat gpurun_out/r2d_probe.jsonl
