O=gpurun_out/s11; mkdir -p $O; cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2_pairs_p2p.json 2> $O/e_p2p.log
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --gather nccl > $O/bench_n2_pairs_nccl.json 2> $O/e_nccl.log
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --mode bands > $O/bench_n2_bands.json 2> $O/e_bands.log
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload 720p_d64_w9 --pairs 32 > $O/bench_n2_720.json 2> $O/e_720.log
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_n2_ref.json 2> $O/e_ref.log
timeout 300 python bench.py --mode bands --steps 10 --warmup 3 > $O/bench_n1_bands.json 2> $O/e_b1.log
nvidia-smi topo -m > $O/topo.txt 2>&1
tail -2 $O/e_*.log; cat $O/*.json
