O=gpurun_out/s10; mkdir -p $O; cd /root/repo
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_ssd.json 2> $O/bench_ssd.err
timeout 300 python bench.py --steps 10 --warmup 3 --cost ncc --no-cpu > $O/bench_ncc.json 2> $O/bench_ncc.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload 1080p_d128_w9 --no-cpu > $O/bench_1080_ssd.json 2>$O/e1
timeout 300 python bench.py --steps 10 --warmup 3 --workload 720p_d64_w9 --pairs 16 --no-cpu > $O/bench_720_ssd.json 2>$O/e2
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_ssd.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_ -c 6 -o $O/full_ssd python tools/profile_one.py 4k_d256_w11 1 ssd > $O/ncu_full_ssd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_ -c 6 -o $O/full_ncc python tools/profile_one.py 4k_d256_w11 1 ncc > $O/ncu_full_ncc.log 2>&1
cat $O/pytest.log; cat $O/smoke.log | tail -3; tail -3 $O/bench_ssd.err; cat $O/bench_ssd.json
