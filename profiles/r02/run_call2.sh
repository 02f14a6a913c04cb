#!/bin/bash
# r02 call 2: persistent multi-level scan kernel + cooperative repair kernel, first run (1 GPU)
O=gpurun_out/c2; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -12 $O/sanity.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scale.py > $O/pytest_old.log 2>&1; echo "old suite rc=$?"; tail -15 $O/pytest_old.log
timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -q --durations=10 > $O/pytest_scale.log 2>&1; echo "scale suite rc=$?"; tail -8 $O/pytest_scale.log
timeout 800 python bench.py --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?"; tail -c 400 $O/bench_1gpu.err
S="--no-cpu-baseline --legs none --no-sustained --steps 20 --warmup 5"
timeout 300 python bench.py $S --opt hybrid=0 --sweep 1,2,8,64,128,256 > $O/bench_hybrid0.json 2> $O/bench_hybrid0.err; echo "hybrid0 rc=$?"
timeout 300 python bench.py $S --sweep 1,2,8,64,128,256 > $O/bench_hybrid1.json 2> $O/bench_hybrid1.err; echo "hybrid1 rc=$?"
timeout 300 python bench.py $S --rows 125000 --sweep 1,4 --opt hybrid=0 > $O/bench_125k_hybrid0.json 2> $O/bench_125k_hybrid0.err
timeout 300 python bench.py $S --rows 125000 --sweep 1,4 > $O/bench_125k_hybrid1.json 2> $O/bench_125k_hybrid1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --legs none --no-sustained > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 900 compute-sanitizer --tool memcheck python tests/sanitizer_check.py > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_memcheck.log
