mkdir -p gpurun_out/s9; cd /root/repo
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/s9/pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s9/bench_ssd.json 2> gpurun_out/s9/bench_ssd.err
python bench.py --steps 10 --warmup 3 --workload 720p_d64_w9 --pairs 16 --no-cpu > gpurun_out/s9/bench_720_ssd.json 2>gpurun_out/s9/e2
cat gpurun_out/s9/pytest.log; tail -3 gpurun_out/s9/bench_ssd.err
