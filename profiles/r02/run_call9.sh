#!/bin/bash
O=gpurun_out/c9; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1,8,64,128,1024 > $O/trace_c2_hybrid1.json 2> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,8 --opt hybrid=0 > $O/trace_c2_hybrid0.json 2>> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,8,1024 --rows 125000 > $O/trace_125k.json 2>> $O/trace.err; echo "rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scale.py > $O/pytest_old.log 2>&1; echo "old suite rc=$?"; tail -n 3 $O/pytest_old.log
tail -n 3 $O/trace.err
