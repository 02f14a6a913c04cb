#!/bin/bash
# r02 call 25: evidence with the final code (PDL chain, finalize changes, limits up to 16384, C4 leg): GPU suite, smoke, both bench
# arms (the driver's default commands), level traces, launch list, ncu full captures of the persistent scan, sanitizer runs
O=gpurun_out/c25; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "full suite rc=$?"; tail -n 4 $O/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 $O/smoke.log
timeout 400 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
SECONDS=0
timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? wall=${SECONDS}s"; tail -c 300 $O/bench_default.err
timeout 300 python profiles/r02/trace_levels.py --batches 1,8,128,1024 > $O/trace_c2.json 2> $O/trace.err; echo "trace rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --legs none --no-sustained > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 8 -c 1 -o $O/prof_scan_b1024 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/ncu_b1024.log 2>&1; echo "ncu b1024 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 8 -c 1 -o $O/prof_scan_b1 python bench.py --batch 1 --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/ncu_b1.log 2>&1; echo "ncu b1 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:finalize_kernel -s 8 -c 1 -o $O/prof_finalize_b1024 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/ncu_fin.log 2>&1; echo "ncu fin rc=$?"
timeout 500 compute-sanitizer --tool memcheck python tests/sanitizer_check.py > $O/sanitizer_memcheck.log 2>&1; tail -n 3 $O/sanitizer_memcheck.log
SAN_CASES=0,1,4 timeout 500 compute-sanitizer --tool racecheck python tests/sanitizer_check.py > $O/sanitizer_racecheck.log 2>&1; tail -n 3 $O/sanitizer_racecheck.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c25/bench_default.json").read().strip().splitlines()[-1])
print("C2", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], d["roofline"]["step_frac"], "parity", d.get("parity_ids_match_oracle"), "b1", d["batch1"]["value"], d["batch1"]["ms_per_step"], "clk", d["clocks"]["sm_mhz"])
for leg,v in d.get("legs",{}).items():
    if "error" in v: print(leg, v); continue
    print(leg, {b:(round(x["qps"],1), round(x["ms_per_step"],3), round(x["e2e"]["value"],1), round(x["roofline"]["frac"],3), round(x["roofline"]["step_frac"],3)) for b,x in v["batches"].items()}, v["parity"].get("parity_ids_match_oracle"), {k:v.get(k) for k in ("wide_rescored_queries","repaired_queries","uncertified_queries","queries")})
PY
ls -la $O
