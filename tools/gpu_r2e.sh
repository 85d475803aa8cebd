#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -3 gpurun_out/r2e_pytest.log
: > gpurun_out/r2e_probe.jsonl
for tag in "" k2lea5 k2lea15 nolea; do
  echo "{\"variant\": \"$tag\"}" >> gpurun_out/r2e_probe.jsonl
  STEREO_LIB_TAG=$tag python tools/probe_hot.py 2160,3840,256,5,ssd,4 >> gpurun_out/r2e_probe.jsonl 2>> gpurun_out/r2e_probe.err
done
echo '{"variant": "sched0"}' >> gpurun_out/r2e_probe.jsonl
STEREO_FAST_SCHED=0 python tools/probe_hot.py 2160,3840,256,5,ssd,4 2160,3840,256,5,ncc,2 1080,1920,128,4,ssd,4 >> gpurun_out/r2e_probe.jsonl 2>> gpurun_out/r2e_probe.err
echo '{"variant": "sched auto"}' >> gpurun_out/r2e_probe.jsonl
python tools/probe_hot.py 2160,3840,256,5,ncc,2 1080,1920,128,4,ssd,4 720,1280,64,4,ssd,4 >> gpurun_out/r2e_probe.jsonl 2>> gpurun_out/r2e_probe.err
for sch in 0 1; do
  STEREO_FAST_SCHED=$sch timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fast_cost_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2e_dram_sched$sch.csv python tools/probe_hot.py 2160,3840,256,5,ssd,4 > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_cost_kernel -s 2 -c 1 -o gpurun_out/r2e_fused_4k python tools/probe_hot.py 2160,3840,256,5,ssd,4 > gpurun_out/r2e_ncu_full.log 2>&1
cat gpurun_out/r2e_probe.jsonl; cat gpurun_out/r2e_dram_sched*.csv | tail -12
