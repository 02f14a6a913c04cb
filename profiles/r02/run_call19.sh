#!/bin/bash
# r02 call 19: finalize fused into the tail of the persistent scan: GPU suite, A/B sweeps (fuse_finalize 1 / 0), traces
O=gpurun_out/c19; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "full suite rc=$?"; tail -n 8 $O/pytest_all.log
for f in 1 0; do
  timeout 600 python bench.py --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,8,64,128,256,1024,4096 --opt fuse_finalize=$f > $O/sweep_c2_fuse$f.json 2> $O/sweep_c2_fuse$f.err; echo "sweep fuse=$f rc=$?"
done
timeout 600 python bench.py --rows 125000 --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,128,1024 > $O/sweep_125k.json 2> $O/sweep_125k.err; echo "sweep 125k rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,128,1024 > $O/trace_c2.json 2> $O/trace.err; echo "trace rc=$?"
