#!/bin/bash
# r02 call 26: box calibration (cuBLAS bf16 / copy on THIS box) beside the bench; same-box A/B of the session's start
# (profiles/r02/variants/libavs_r02base.so = commit 01feaa7) vs the current library with pdl 1 / 0; default bench with the C1 leg
O=gpurun_out/c26; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu,power.draw --format=csv > $O/gpu.txt
S="--no-cpu-baseline --legs none --no-sustained --steps 30 --warmup 5"
for rep in 1 2; do
  AVS_LIB=$PWD/profiles/r02/variants/libavs_r02base.so timeout 300 python bench.py $S --sweep 1,128,1024 > $O/c2_r02base_$rep.json 2> $O/c2_r02base_$rep.err; echo "r02base $rep rc=$?"
  timeout 300 python bench.py $S --sweep 1,128,1024 --opt pdl=1 > $O/c2_pdl1_$rep.json 2> $O/c2_pdl1_$rep.err; echo "pdl1 $rep rc=$?"
  timeout 300 python bench.py $S --sweep 1,128,1024 --opt pdl=0 > $O/c2_pdl0_$rep.json 2> $O/c2_pdl0_$rep.err; echo "pdl0 $rep rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c26/c2_*_[12].json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], [(x["batch"], round(x["ms_per_step"],4), round(x["batch"]/x["e2e"]*1e3,4), round(x["kernel_ms"],4)) for x in d.get("sweep",[])], d.get("box_calibration"))
    except Exception as e: print(f, "ERR", e)
PY
SECONDS=0
timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? wall=${SECONDS}s"; tail -c 300 $O/bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c26/bench_default.json").read().strip().splitlines()[-1])
print("C2", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], d["roofline"]["step_frac"], "parity", d.get("parity_ids_match_oracle"), "b1", d["batch1"]["value"], d["batch1"]["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d.get("box_calibration"))
for leg,v in d.get("legs",{}).items():
    if not v or "error" in v or "batches" not in v: print(leg, v); continue
    print(leg, {b:(round(x["qps"],1), round(x["ms_per_step"],3), round(x["e2e"]["value"],1), round(x["roofline"]["frac"],3), round(x["roofline"]["step_frac"],3)) for b,x in v["batches"].items()}, v["parity"].get("parity_ids_match_oracle"), {k:v.get(k) for k in ("wide_rescored_queries","repaired_queries","uncertified_queries","queries")})
PY
