#!/bin/bash
# r02 call 16: sorting-network boot epilogue + heads-based warp select: parity (boot forced for every batch), traces, A/B
O=gpurun_out/c16; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "boot_level" > $O/pytest_boot.log 2>&1; echo "boot tests rc=$?"; tail -n 5 $O/pytest_boot.log
timeout 300 python profiles/r02/trace_levels.py --batches 1,64,128,256,1024 --opt boot=2 > $O/trace_c2_boot2.json 2> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,128,1024 --rows 125000 --opt boot=2 > $O/trace_125k_boot2.json 2>> $O/trace.err; echo "rc=$?"
tail -n 3 $O/trace.err
for b in 2 1; do
  timeout 600 python bench.py --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,8,64,128,256,512,1024,4096 --opt boot=$b > $O/sweep_c2_boot$b.json 2> $O/sweep_c2_boot$b.err; echo "sweep boot=$b rc=$?"
done
timeout 600 python bench.py --rows 125000 --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,128,1024 --opt boot=2 > $O/sweep_125k_boot2.json 2> $O/sweep_125k_boot2.err; echo "sweep 125k boot=2 rc=$?"
timeout 600 python bench.py --rows 125000 --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,128,1024 --opt boot=1 > $O/sweep_125k_boot1.json 2> $O/sweep_125k_boot1.err; echo "sweep 125k boot=1 rc=$?"
