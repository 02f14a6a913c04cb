#!/bin/bash
# r02 call 27 (the round's last GPU minutes, ordered by value): the TMA producer's nested tile loop restored (the flat cursor
# of calls 21-26 cost the compute-bound path 30 %): same-box A/B vs the session's start (libavs_r02base.so = commit 01feaa7),
# the driver's default bench command, launch list + ncu capture of the final code, then the GPU suite with what is left.
O=gpurun_out/c27; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu,power.draw --format=csv > $O/gpu.txt
S="--no-cpu-baseline --legs none --no-sustained --steps 30 --warmup 5"
AVS_LIB=$PWD/profiles/r02/variants/libavs_r02base.so timeout 200 python bench.py $S --sweep 1,128,1024 > $O/c2_r02base_1.json 2> $O/c2_r02base_1.err; echo "r02base rc=$?"
timeout 200 python bench.py $S --sweep 1,128,1024 > $O/c2_cur_1.json 2> $O/c2_cur_1.err; echo "cur rc=$?"
timeout 120 python bench.py $S --rows 125000 --sweep 1,128,1024 > $O/s125k_cur_1.json 2> $O/s125k_cur_1.err; echo "125k cur rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c27/*_1.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], [(x["batch"], round(x["ms_per_step"],4), round(x["batch"]/x["e2e"]*1e3,4), round(x["kernel_ms"],4)) for x in d.get("sweep",[])])
    except Exception as e: print(f, "ERR", e)
PY
SECONDS=0
timeout 480 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? wall=${SECONDS}s"; tail -c 300 $O/bench_default.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c27/bench_default.json").read().strip().splitlines()[-1])
    print("C2", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], d["roofline"]["step_frac"], "parity", d.get("parity_ids_match_oracle"), "b1", d["batch1"]["value"], d["batch1"]["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d.get("box_calibration"))
    for leg,v in d.get("legs",{}).items():
        if not v or "error" in v or "batches" not in v: print(leg, v); continue
        print(leg, {b:(round(x["qps"],1), round(x["ms_per_step"],3), round(x["e2e"]["value"],1), round(x["roofline"]["frac"],3), round(x["roofline"]["step_frac"],3)) for b,x in v["batches"].items()}, v["parity"].get("parity_ids_match_oracle"), {k:v.get(k) for k in ("wide_rescored_queries","repaired_queries","uncertified_queries","queries")})
except Exception as e: print("bench line ERR", e)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --legs none --no-sustained > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 8 -c 1 -o $O/prof_scan_b1024 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/ncu_b1024.log 2>&1; echo "ncu b1024 rc=$?"
SECONDS=0
timeout 420 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "full suite rc=$? wall=${SECONDS}s"; tail -n 4 $O/pytest_all.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 $O/smoke.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 8 -c 1 -o $O/prof_scan_b1 python bench.py --batch 1 --steps 3 --warmup 2 --no-cpu-baseline --legs none --no-sustained --only-batch > $O/ncu_b1.log 2>&1; echo "ncu b1 rc=$?"
