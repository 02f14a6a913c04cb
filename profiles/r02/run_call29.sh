#!/bin/bash
# r02 call 29 (the round's last GPU minutes): the eps rule in the warp's radix select (levels that collect > 256 keys) -
# C4 shard as the main workload (call 28: repair_kernel 25.6 ms in EVERY batch-1024 search), then the GPU suite with the
# new test first, then the driver's default bench command if time is left.
O=gpurun_out/c29; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu,power.draw --format=csv > $O/gpu.txt
timeout 90 python bench.py --rows 2500000 --dim 3072 --k 50 --steps 10 --warmup 3 --no-cpu-baseline --legs none --no-sustained > $O/bench_c4_main.json 2> $O/bench_c4_main.err; echo "c4 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c29/bench_c4_main.json").read().strip().splitlines()[-1])
    print("C4 main", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], d["roofline"]["step_frac"], "parity", d.get("parity_ids_match_oracle"), "b1", d["batch1"]["ms_per_step"], {k:d.get(k) for k in ("wide_rescored_queries","repaired_queries","uncertified_queries","queries")})
except Exception as e: print("ERR", e)
PY
SECONDS=0
timeout 300 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_all.log 2>&1; echo "suite rc=$? wall=${SECONDS}s"; tail -n 5 $O/pytest_all.log
SECONDS=0
timeout 100 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c29/bench_default.json").read().strip().splitlines()[-1])
    print("C2", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], d["roofline"]["step_frac"], "parity", d.get("parity_ids_match_oracle"), "b1", d["batch1"]["value"], d["batch1"]["ms_per_step"])
    for leg,v in d.get("legs",{}).items():
        if not v or "error" in v or "batches" not in v: print(leg, str(v)[:300]); continue
        print(leg, {b:(round(x["qps"],1), round(x["ms_per_step"],3), round(x["e2e"]["value"],1), round(x["roofline"]["frac"],3), round(x["roofline"]["step_frac"],3)) for b,x in v["batches"].items()}, v["parity"].get("parity_ids_match_oracle"), {k:v.get(k) for k in ("wide_rescored_queries","repaired_queries","uncertified_queries","queries")})
except Exception as e: print("bench line ERR", e)
PY
