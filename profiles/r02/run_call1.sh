#!/bin/bash
# r02 call 1: baseline of the round-1 kernels under the new tests / bench legs (1 GPU)
mkdir -p gpurun_out/c1
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1/gpu.txt 2>&1
free -g > gpurun_out/c1/host_mem.txt; nproc >> gpurun_out/c1/host_mem.txt
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scale.py > gpurun_out/c1/pytest_old.log 2>&1; echo "old suite rc=$?"
timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -q --durations=10 > gpurun_out/c1/pytest_scale.log 2>&1; echo "scale suite rc=$?"
tail -5 gpurun_out/c1/pytest_scale.log
timeout 800 python bench.py --steps 20 --warmup 5 > gpurun_out/c1/bench_1gpu.json 2> gpurun_out/c1/bench_1gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/c1/bench_1gpu.err
