#!/bin/bash
O=gpurun_out/c7; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1,8,128,1024 > $O/trace_c2.json 2> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1024 --opt gemm_dense_rows=1024 > $O/trace_c2_dense1024.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1024 --opt gemm_dense_rows=512 > $O/trace_c2_dense512.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1024 --opt final_sigma=3 > $O/trace_c2_sigma3.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1024 --opt fine_ratio=8 > $O/trace_c2_fine8.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1024 --opt fine_ratio=2 > $O/trace_c2_fine2.json 2>> $O/trace.err; echo "rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scale.py > $O/pytest_old.log 2>&1; echo "old suite rc=$?"; tail -n 3 $O/pytest_old.log
timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -q > $O/pytest_scale.log 2>&1; echo "scale suite rc=$?"; tail -n 3 $O/pytest_scale.log
S="--no-cpu-baseline --legs none --no-sustained --steps 20 --warmup 5"
timeout 300 python bench.py $S --sweep 1,2,8,64,128,256,512,2048,4096 > $O/bench_sweep.json 2> $O/bench_sweep.err; echo "sweep rc=$?"
timeout 300 python bench.py $S --rows 12500000 --k 100 --sweep 1,1024 > $O/bench_c5shard.json 2> $O/bench_c5shard.err; echo "c5 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --legs none --no-sustained > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -n 3 $O/trace.err
