#!/bin/bash
# r02 call 12: 2 x B200 - multi-rank parity (fused peer-memory exchange, NCCL fallback, host-buffer call) and bench --gpus 2
O=gpurun_out/c12; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > $O/dist_check_2gpu.log 2>&1; echo "dist check rc=$?"; grep -E "world=|Error|error" $O/dist_check_2gpu.log | tail -n 12
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -k "multi_gpu or c2_full" > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -n 3 $O/pytest_multi.log
timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"; tail -c 600 $O/bench_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-p2p --legs none --no-sustained > $O/bench_2gpu_nccl.json 2> $O/bench_2gpu_nccl.err; echo "bench2 nccl rc=$?"
