#!/bin/bash
# r02 call 14: boot level (per-thread top-J lists as the threshold-free level of the tensor-core scan): parity, traces, A/B
O=gpurun_out/c14; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "boot_level" > $O/pytest_boot.log 2>&1; echo "boot tests rc=$?"; tail -n 15 $O/pytest_boot.log
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1,8,64,128,1024 > $O/trace_c2_boot1.json 2> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,8,64,128,1024 --opt boot=0 > $O/trace_c2_boot0.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,8,1024 --rows 125000 > $O/trace_125k_boot1.json 2>> $O/trace.err; echo "rc=$?"
tail -n 3 $O/trace.err
for b in 1 0; do
  timeout 600 python bench.py --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,2,8,16,64,128,256,1024 --opt boot=$b > $O/sweep_c2_boot$b.json 2> $O/sweep_c2_boot$b.err; echo "sweep boot=$b rc=$?"
done
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "full suite rc=$?"; tail -n 5 $O/pytest_all.log
