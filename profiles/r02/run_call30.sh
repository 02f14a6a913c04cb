#!/bin/bash
# r02 call 30 (the last GPU minute): launch list of the C4 shard with the eps rule in the radix select (compare c28/launches_c4.csv)
O=gpurun_out/c30; mkdir -p $O
timeout 50 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c4.csv python bench.py --rows 2500000 --dim 3072 --k 50 --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/bench_c4_under_ncu.json 2> $O/ncu_c4.err; echo "c4 launch list rc=$?"
