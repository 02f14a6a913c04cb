#!/bin/bash
# r02 call 15: boot level for HBM-bound batches + lean CTA select: parity, traces with select phase stamps, A/B sweep
O=gpurun_out/c15; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "boot_level" > $O/pytest_boot.log 2>&1; echo "boot tests rc=$?"; tail -n 5 $O/pytest_boot.log
timeout 300 python profiles/r02/trace_levels.py --batches 1,8,64,128 > $O/trace_c2.json 2> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,8,128 --rows 125000 > $O/trace_125k.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,8,128 --rows 10000000 --dim 1024 --metric IP > $O/trace_c3.json 2>> $O/trace.err; echo "rc=$?"
tail -n 3 $O/trace.err
timeout 600 python bench.py --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,2,8,16,64,128,256,1024 > $O/sweep_c2.json 2> $O/sweep_c2.err; echo "sweep rc=$?"
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "full suite rc=$?"; tail -n 5 $O/pytest_all.log
