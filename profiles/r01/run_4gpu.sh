# 4 x B200 with the final round-1 code: sharded parity (both exchange modes), C2 strong-scaling line, C5 (100 M x 768, top-100) on 4 GPUs
O=gpurun_out/r01l; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/dist_gpu_check.py > $O/dist_check_4gpu.log 2>&1; grep -E "OK|MISMATCH" $O/dist_check_4gpu.log
timeout 200 $TR bench.py --gpus 4 --steps 50 --warmup 5 > $O/bench_4gpu_c2.json 2> $O/bench_4gpu_c2.err
timeout 400 $TR bench.py --gpus 4 --rows 100000000 --k 100 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_4gpu_c5.json 2> $O/bench_4gpu_c5.err
tail -c 300 $O/bench_4gpu_c5.err
