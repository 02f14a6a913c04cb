#!/bin/bash
# r02 call 3: rank-counting select; phase traces; fp8 feasibility measurement (1 GPU)
O=gpurun_out/c3; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -3 $O/sanity.log
timeout 300 python profiles/r02/trace_levels.py > $O/trace_c2_hybrid1.json 2> $O/trace.err; echo "trace rc=$?"
timeout 300 python profiles/r02/trace_levels.py --opt hybrid=0 > $O/trace_c2_hybrid0.json 2>> $O/trace.err
timeout 300 python profiles/r02/trace_levels.py --rows 125000 --opt hybrid=0 > $O/trace_125k_hybrid0.json 2>> $O/trace.err
timeout 300 python profiles/r02/trace_levels.py --rows 125000 > $O/trace_125k_hybrid1.json 2>> $O/trace.err
timeout 300 python profiles/r02/trace_levels.py --opt fine_ratio=8 --batches 1024 > $O/trace_c2_fine8.json 2>> $O/trace.err
timeout 300 python profiles/r02/trace_levels.py --opt final_sigma=3 --batches 1024 > $O/trace_c2_sigma3.json 2>> $O/trace.err
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scale.py > $O/pytest_old.log 2>&1; echo "old suite rc=$?"; tail -3 $O/pytest_old.log
timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -q > $O/pytest_scale.log 2>&1; echo "scale suite rc=$?"; tail -3 $O/pytest_scale.log
S="--no-cpu-baseline --legs none --no-sustained --steps 20 --warmup 5"
timeout 300 python bench.py $S --opt hybrid=0 --sweep 1,2,8,64,128,256 > $O/bench_hybrid0.json 2> $O/bench_hybrid0.err; echo "hybrid0 rc=$?"
timeout 300 python bench.py $S --sweep 1,2,8,64,128,256 > $O/bench_hybrid1.json 2> $O/bench_hybrid1.err; echo "hybrid1 rc=$?"
timeout 300 python profiles/r02/fp8_feasibility.py --rows 1000000 --dim 768 --k 10 > $O/fp8_c2.json 2> $O/fp8.err; echo "fp8 c2 rc=$?"
timeout 300 python profiles/r02/fp8_feasibility.py --rows 12500000 --dim 768 --k 100 > $O/fp8_c5shard.json 2>> $O/fp8.err; echo "fp8 c5 rc=$?"
timeout 300 python profiles/r02/fp8_feasibility.py --rows 2500000 --dim 3072 --k 50 --nq 32 > $O/fp8_c4shard.json 2>> $O/fp8.err; echo "fp8 c4 rc=$?"
tail -3 $O/trace.err $O/fp8.err
