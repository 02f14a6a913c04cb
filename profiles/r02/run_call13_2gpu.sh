#!/bin/bash
# r02 call 13: 2 x B200 - bench --gpus 2 after the wall-clock exchange timeout + barrier fix (call 12 died in the sustained leg:
# rank 1 entered a collective search while rank 0 was still in the host-side oracle check, and its spin-count bound expired)
O=gpurun_out/c13; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"; tail -c 600 $O/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/dist_gpu_check.py > $O/dist_check_2gpu.log 2>&1; echo "dist check rc=$?"; grep -E "world=|Error|error" $O/dist_check_2gpu.log | tail -n 8
