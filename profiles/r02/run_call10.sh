#!/bin/bash
# r02 call 10: full default bench (1 GPU) + reference arm + ncu full capture of the persistent scan kernel
O=gpurun_out/c10; mkdir -p $O
timeout 800 python bench.py --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?"; tail -c 300 $O/bench_1gpu.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
# dominant kernel, batch 1024 (CTA pairs): skip the fill / warm-up launches, capture one persistent scan launch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 8 -c 1 -o $O/prof_scan_b1024 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/ncu_b1024.log 2>&1; echo "ncu b1024 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 8 -c 1 -o $O/prof_scan_b1 python bench.py --batch 1 --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/ncu_b1.log 2>&1; echo "ncu b1 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:finalize_kernel -s 8 -c 1 -o $O/prof_finalize_b1024 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/ncu_fin.log 2>&1; echo "ncu fin rc=$?"
ls -la $O
