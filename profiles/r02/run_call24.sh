#!/bin/bash
# r02 call 24: two-level schedule for small stores on the compute-bound path (option boot2_ratio: boot level directly in front
# of the final level): same-box A/B on 125 K / 250 K / 500 K-row shards and C2, batch 256 / 1024 / 4096; parity tests with it on
O=gpurun_out/c24; mkdir -p $O
S="--no-cpu-baseline --legs none --no-sustained --steps 30 --warmup 5"
for rep in 1 2; do
for r in 0 12 16 32; do
  for rows in 125000 250000 500000; do
    timeout 300 python bench.py $S --rows $rows --sweep 256,1024,4096 --opt boot2_ratio=$r > $O/s${rows}_r${r}_$rep.json 2> $O/s${rows}_r${r}_$rep.err; echo "rows $rows r=$r rep $rep rc=$?"
  done
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c24/s*_[12].json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], [(x["batch"], x["levels"], round(x["ms_per_step"],4), round(x["kernel_ms"],4)) for x in d.get("sweep",[])], d.get("parity_ids_match_oracle"))
    except Exception as e: print(f, "ERR", e)
PY
AVS_OPTS="boot2_ratio=16" timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x > $O/pytest_boot2.log 2>&1; echo "tests with boot2_ratio=16 rc=$?"; tail -n 4 $O/pytest_boot2.log
