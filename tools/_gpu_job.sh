mkdir -p gpurun_out/s8; cd /root/repo
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/s8/pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s8/bench_ssd.json 2> gpurun_out/s8/bench_ssd.err
python bench.py --steps 10 --warmup 3 --cost ncc --no-cpu > gpurun_out/s8/bench_ncc.json 2> gpurun_out/s8/bench_ncc.err
python bench.py --steps 10 --warmup 3 --workload 1080p_d128_w9 --no-cpu > gpurun_out/s8/bench_1080_ssd.json 2>gpurun_out/s8/e1
python bench.py --steps 10 --warmup 3 --workload 720p_d64_w9 --pairs 16 --no-cpu > gpurun_out/s8/bench_720_ssd.json 2>gpurun_out/s8/e2
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s8/launches_ssd.csv python tools/profile_one.py 4k_d256_w11 8 ssd > gpurun_out/s8/p1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s8/launches_720.csv python tools/profile_one.py 720p_d64_w9 8 ssd > gpurun_out/s8/p2.log 2>&1
cat gpurun_out/s8/pytest.log
