#!/bin/bash
# r02 call 18: evidence with the boot-level code: GPU suite, both bench arms (default command), launch list, ncu full captures
O=gpurun_out/c18; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "full suite rc=$?"; tail -n 4 $O/pytest_all.log
timeout 400 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 300 $O/bench_default.err
timeout 300 python profiles/r02/trace_levels.py --batches 1,8,128,1024 > $O/trace_c2.json 2> $O/trace.err; echo "trace rc=$?"
# launch list of the default command (short): every kernel of the timed region with its duration
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --legs none --no-sustained > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 8 -c 1 -o $O/prof_scan_b1024 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/ncu_b1024.log 2>&1; echo "ncu b1024 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 8 -c 1 -o $O/prof_scan_b1 python bench.py --batch 1 --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/ncu_b1.log 2>&1; echo "ncu b1 rc=$?"
timeout 500 compute-sanitizer --tool memcheck python tests/sanitizer_check.py > $O/sanitizer_memcheck.log 2>&1; tail -n 3 $O/sanitizer_memcheck.log
ls -la $O
