#!/bin/bash
# r02 call 20: programmatic dependent launch between the kernels of a search, candidate pruning in finalize, limits above 256
# (exact master scan): new tests first, then the full GPU suite, then same-box A/B sweeps (pdl 1 / 0) on C2 and a 125 K-row shard
O=gpurun_out/c20; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 300 python tests/sanitizer_check.py > $O/sanity.log 2>&1; echo "sanity rc=$?"; tail -n 2 $O/sanity.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "limits_above or large_limit or programmatic or rescoring_skips" > $O/pytest_new.log 2>&1; echo "new tests rc=$?"; tail -n 15 $O/pytest_new.log
for p in 1 0; do
  timeout 600 python bench.py --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,8,128,256,1024,4096 --opt pdl=$p > $O/sweep_c2_pdl$p.json 2> $O/sweep_c2_pdl$p.err; echo "sweep c2 pdl=$p rc=$?"
  timeout 600 python bench.py --rows 125000 --steps 20 --warmup 5 --legs none --no-sustained --no-cpu-baseline --sweep 1,128,1024 --opt pdl=$p > $O/sweep_125k_pdl$p.json 2> $O/sweep_125k_pdl$p.err; echo "sweep 125k pdl=$p rc=$?"
done
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "full suite rc=$?"; tail -n 8 $O/pytest_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c20/sweep_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], [(x["batch"], round(x["ms_per_step"],4), round(x["batch"]/x["e2e"]*1e3,4), round(x["kernel_ms"],4)) for x in d.get("sweep",[])])
    except Exception as e: print(f, "ERR", e)
PY
