# coalesced dense-level stores in the tensor-core scan: tests, racecheck of a tensor-core case, timing, launch list
O=gpurun_out/r01o; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
SAN_CASES=1,5,8 timeout 300 compute-sanitizer --tool racecheck python tests/sanitizer_check.py > $O/sanitizer_racecheck_gemm.log 2>&1; grep -E "OK|MISMATCH|RACECHECK SUMMARY" $O/sanitizer_racecheck_gemm.log
python bench.py --steps 30 > $O/bench_default.json 2> $O/bench_default.err; cut -c1-140 $O/bench_default.json
python bench.py --no-cpu-baseline --only-batch --sweep 16,64,128,256,1024 --steps 30 --warmup 5 > $O/sweep_mid.json 2> $O/sweep_mid.err
python bench.py --no-cpu-baseline --only-batch --rows 125000 --sweep 64,1024 --steps 30 --warmup 5 > $O/shard125k.json 2> $O/shard125k.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
