#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
python tools/probe_hot.py 128,640,96,7,ssd,1 256,640,96,7,ssd,1 511,640,96,7,ssd,1 1022,640,96,7,ssd,1 2044,640,96,7,ssd,1 4088,640,96,7,ssd,1 \
   511,640,96,7,ncc,1 2044,640,96,7,ncc,1 511,640,96,7,ssd,1,noisy 2044,640,96,7,ssd,1,noisy 511,640,96,7,ncc,1,noisy 2044,640,96,7,ncc,1,noisy \
   511,640,96,7,ssd,1,f32 128,128,4,6,ssd,1,f32 511,640,128,5,ssd,1 2044,640,128,5,ssd,1 720,1280,64,4,ssd,4 720,1280,64,4,ssd,1 1080,1920,128,4,ssd,4 \
   2160,3840,256,5,ssd,4 > gpurun_out/r2c_probe.jsonl 2> gpurun_out/r2c_probe.err
cat gpurun_out/r2c_probe.jsonl
