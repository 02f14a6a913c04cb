# Round-1 evidence run on one B200 (final code): GPU tests, both bench arms, batch sweeps, per-GPU shards of the
# 8-GPU configs, the ncu launch list of the default bench command, compute-sanitizer.
set -x
O=gpurun_out/r01g; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench_default.json 2> $O/bench_default.err
S="--no-cpu-baseline --only-batch"
python bench.py $S --sweep 1,2,3,4,8,16,32,64,128,256,512,1024,2048,4096 --steps 20 --warmup 3 > $O/sweep_c2.json 2> $O/sweep_c2.err
python bench.py $S --rows 10000000 --dim 1024 --metric IP --sweep 1,2,4,8,16,32,64,128,256,512,1024,2048,4096 --steps 10 --warmup 3 > $O/sweep_c3.json 2> $O/sweep_c3.err
python bench.py $S --rows 12500000 --dim 768 --k 100 --sweep 1,64,1024 --steps 10 --warmup 3 > $O/shard_c5.json 2> $O/shard_c5.err
python bench.py $S --rows 2500000 --dim 3072 --k 50 --sweep 1,64,1024 --steps 10 --warmup 3 > $O/shard_c4.json 2> $O/shard_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
timeout 500 compute-sanitizer --tool memcheck python tests/sanitizer_check.py > $O/sanitizer_memcheck.log 2>&1; tail -3 $O/sanitizer_memcheck.log
tail -3 $O/pytest_gpu.log; cut -c1-300 $O/bench_default.json
