#!/bin/bash
# r02 call 28 (diagnostic, ~1 min): where do the 29 ms outside the scan go in the C4-shard batch-1024 step (2.5 M x 3072, top-50:
# scan 11.0 ms, step 39.9 ms in c27/bench_default.json)?  Launch list of that configuration as the main workload.
O=gpurun_out/c28; mkdir -p $O
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv python bench.py --rows 2500000 --dim 3072 --k 50 --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/bench_c4_under_ncu.json 2> $O/ncu_c4.err; echo "c4 launch list rc=$?"
grep -c . $O/launches_c4.csv
