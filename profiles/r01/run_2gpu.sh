# 2 x B200: sharded parity (both exchange modes) and the strong-scaling bench line; C2/8-sized shards (125 K rows per GPU)
O=gpurun_out/r01f; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR tests/dist_gpu_check.py > $O/dist_check_2gpu.log 2>&1; tail -8 $O/dist_check_2gpu.log
timeout 300 $TR bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
timeout 300 $TR bench.py --gpus 2 --rows 250000 --steps 50 --warmup 5 > $O/bench_2gpu_125k_shards.json 2> $O/bench_2gpu_125k_shards.err
