#!/bin/bash
# r02 call 23: 2 x B200 with the final code (PDL chain, finalize changes): multi-GPU parity check, the pytest torchrun test,
# and the driver's bench command at N = 2
O=gpurun_out/c23; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/dist_gpu_check.py > $O/dist_check_2gpu.log 2>&1; echo "dist check rc=$?"; grep -E "world=|Error|error" $O/dist_check_2gpu.log | tail -n 8
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k torchrun > $O/pytest_torchrun.log 2>&1; echo "pytest torchrun rc=$?"; tail -n 3 $O/pytest_torchrun.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"; tail -c 400 $O/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c23/bench_2gpu.json").read().strip().splitlines()[-1])
print("C2 N=2", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d.get("parity_ids_match_oracle"), "exchange_us", d["engine"].get("exchange_kernel_us"), "b1", d.get("batch1",{}).get("value"))
for leg,v in d.get("legs",{}).items():
    print(leg, {b:(round(x["qps"],1), round(x["ms_per_step"],3), round(x["e2e"]["value"],1)) for b,x in v["batches"].items()}, v["parity"].get("parity_ids_match_oracle"))
PY
