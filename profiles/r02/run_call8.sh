#!/bin/bash
O=gpurun_out/c8; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1,2,8,64,128,1024 > $O/trace_c2_hybrid1.json 2> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1,2,8,64,128 --opt hybrid=0 > $O/trace_c2_hybrid0.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,8,1024 --rows 125000 --opt hybrid=0 > $O/trace_125k_hybrid0.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,8 --rows 125000 > $O/trace_125k_hybrid1.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1024 --opt finalize_threads=1024 > $O/trace_c2_fin1024.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1024 --opt finalize_threads=512 > $O/trace_c2_fin512.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,8 --rows 12500000 --k 100 --opt hybrid=0 > $O/trace_c5_hybrid0.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,8 --rows 12500000 --k 100 > $O/trace_c5_hybrid1.json 2>> $O/trace.err; echo "rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scale.py > $O/pytest_old.log 2>&1; echo "old suite rc=$?"; tail -n 3 $O/pytest_old.log
timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -q > $O/pytest_scale.log 2>&1; echo "scale suite rc=$?"; tail -n 3 $O/pytest_scale.log
timeout 900 compute-sanitizer --tool memcheck python tests/sanitizer_check.py > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 2 $O/sanitizer_memcheck.log
tail -n 3 $O/trace.err
