#!/bin/bash
# r02 call 17: boot level for every batch size (insertion variant <= 8 queries, sorting networks above, ~4 tiles per CTA pair on
# the compute-bound schedule): full GPU suite, sweeps on C2 and on a 125 K-row shard, traces
O=gpurun_out/c17; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 300 python profiles/r02/trace_levels.py --batches 1,8,64,256,1024 > $O/trace_c2.json 2> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,128,1024 --rows 125000 > $O/trace_125k.json 2>> $O/trace.err; echo "rc=$?"
tail -n 3 $O/trace.err
timeout 600 python bench.py --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,8,64,128,256,512,1024,4096 > $O/sweep_c2.json 2> $O/sweep_c2.err; echo "sweep rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,8,64,128,256,512,1024,4096 --opt boot=0 > $O/sweep_c2_boot0.json 2> $O/sweep_c2_boot0.err; echo "sweep boot=0 rc=$?"
timeout 600 python bench.py --rows 125000 --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,128,1024 > $O/sweep_125k.json 2> $O/sweep_125k.err; echo "sweep 125k rc=$?"
timeout 600 python bench.py --rows 125000 --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,128,1024 --opt boot=0 > $O/sweep_125k_boot0.json 2> $O/sweep_125k_boot0.err; echo "sweep 125k boot=0 rc=$?"
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "full suite rc=$?"; tail -n 5 $O/pytest_all.log
