#!/bin/bash
# One N=1 GPU session (run under tools/gpurun.sh): parity tests, smoke, the default bench line (with suite and parity_check),
# the reference arm, the other bench lines, the ncu launch list of the bench and one --set full capture of the hot kernel.
#   bash tools/gpu_session.sh <tag>     -> gpurun_out/<tag>_*
T=${1:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest_gpu.log; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_err0.log; echo "reference arm rc=$?"
timeout 900 python bench.py > gpurun_out/${T}_bench_n1_4k_ssd.json 2> gpurun_out/${T}_err1.log; echo "bench rc=$?"
timeout 600 python bench.py --cost ncc --pairs 2 --no-suite > gpurun_out/${T}_bench_n1_4k_ncc.json 2> gpurun_out/${T}_err2.log; echo "bench ncc rc=$?"
timeout 600 python bench.py --workload 720p_d64_w9 --pairs 16 --no-suite > gpurun_out/${T}_bench_n1_720p_x16_ssd.json 2> gpurun_out/${T}_err3.log; echo "bench 720p rc=$?"
timeout 600 python bench.py --mode bands > gpurun_out/${T}_bench_n1_bands.json 2> gpurun_out/${T}_err4.log; echo "bench bands rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-suite --no-parity > gpurun_out/${T}_launches_bench.log 2>&1; echo "ncu bench launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_cost_kernel -s 2 -c 1 -o gpurun_out/${T}_fused_4k python tools/probe_hot.py 2160,3840,256,5,ssd,4 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/probe_hot.py 511,640,96,7,ssd,1,f32 511,640,96,7,ssd,1,noisy 511,640,96,7,ncc,1,f32 511,640,96,7,ncc,1,noisy 720,1280,64,4,ssd,4 2160,3840,256,5,ssd,1,noisy > gpurun_out/${T}_probe.jsonl 2> gpurun_out/${T}_probe.err; cat gpurun_out/${T}_probe.jsonl
for f in gpurun_out/${T}_err*.log; do tail -n 2 "$f"; done; true
