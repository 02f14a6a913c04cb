#!/bin/bash
# r02 call 22: finalize micro-optimisations (independent loads first, L2 prefetch of the candidate rows, ids in flight with the rows,
# rank-count sort for K' <= 64) vs the previous build (profiles/r02/variants/libavs_c20.so) on ONE box; new-code tests first
O=gpurun_out/c22; mkdir -p $O
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_parity.log 2>&1; echo "parity tests rc=$?"; tail -n 5 $O/pytest_parity.log
S="--no-cpu-baseline --legs none --no-sustained --steps 30 --warmup 5"
for rep in 1 2; do
for v in c20 cur; do
  if [ $v = c20 ]; then export AVS_LIB=$PWD/profiles/r02/variants/libavs_c20.so; else unset AVS_LIB; fi
  timeout 300 python bench.py $S --sweep 1,8,128,256,1024 > $O/c2_${v}_$rep.json 2> $O/c2_${v}_$rep.err; echo "c2 $v $rep rc=$?"
  timeout 300 python bench.py $S --rows 125000 --sweep 1,128,1024 > $O/shard125k_${v}_$rep.json 2> $O/shard125k_${v}_$rep.err; echo "125k $v $rep rc=$?"
  timeout 300 python bench.py $S --rows 2000000 --k 100 --sweep 1,1024 > $O/k100_${v}_$rep.json 2> $O/k100_${v}_$rep.err; echo "k100 $v $rep rc=$?"
done
done
unset AVS_LIB
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c22/*_[12].json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], [(x["batch"], round(x["ms_per_step"],4), round(x["batch"]/x["e2e"]*1e3,4), round(x["kernel_ms"],4)) for x in d.get("sweep",[])])
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --legs none --no-sustained > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
grep -v "^==" $O/launches_default.csv | python -c "
import csv,sys,collections
r=csv.DictReader(sys.stdin); agg=collections.defaultdict(list)
for row in r:
    agg[row['Kernel Name'][:40]].append(float(row['Metric Value']))
for k,v in agg.items(): print(k, len(v), 'median_us', sorted(v)[len(v)//2]/1000, 'min', min(v)/1000, 'max', max(v)/1000)
"
