#!/bin/bash
# r02 call 21: flat TMA-producer loop with the L2 prefetch option: full GPU suite, same-box A/B sweeps (l2_prefetch 0 / 8 / 16 / 32)
# on C2 and a 125 K-row shard (small batches), level traces, launch list of the default command
O=gpurun_out/c21; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "full suite rc=$?"; tail -n 5 $O/pytest_all.log
for p in 0 8 16 32 0; do
  timeout 600 python bench.py --steps 30 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,8,128 --opt l2_prefetch=$p > $O/sweep_c2_pf$p.json 2> $O/sweep_c2_pf$p.err; echo "sweep c2 pf=$p rc=$?"
  timeout 600 python bench.py --rows 125000 --steps 30 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,128 --opt l2_prefetch=$p > $O/sweep_125k_pf$p.json 2> $O/sweep_125k_pf$p.err; echo "sweep 125k pf=$p rc=$?"
  python - <<PY
import json
for f in ["sweep_c2_pf$p.json","sweep_125k_pf$p.json"]:
    try:
        d=json.loads(open("$O/"+f).read().strip().splitlines()[-1])
        print(f, [(x["batch"], round(x["ms_per_step"],4), round(x["batch"]/x["e2e"]*1e3,4), round(x["kernel_ms"],4)) for x in d.get("sweep",[])])
    except Exception as e: print(f, "ERR", e)
PY
done
timeout 300 python profiles/r02/trace_levels.py --batches 1,8,128,1024 > $O/trace_c2.json 2> $O/trace.err; echo "trace rc=$?"
timeout 300 python profiles/r02/trace_levels.py --batches 1,128,1024 --rows 125000 > $O/trace_125k.json 2>> $O/trace.err; echo "trace rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --legs none --no-sustained > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
grep -v "^==" $O/launches_default.csv | python -c "
import csv,sys,collections
r=csv.DictReader(sys.stdin); agg=collections.defaultdict(list)
for row in r:
    agg[row['Kernel Name'][:40]].append(float(row['Metric Value']))
for k,v in agg.items(): print(k, len(v), 'median_us', sorted(v)[len(v)//2]/1000, 'min', min(v)/1000, 'max', max(v)/1000)
"
