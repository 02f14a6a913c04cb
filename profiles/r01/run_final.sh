# final round-1 check of the hybrid policy: GPU tests, default bench, C2 sweep, small-batch rows of the big configs,
# launch list of the default bench command, one full ncu capture of the batch-1 final-level scan
O=gpurun_out/r01i; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err
S="--no-cpu-baseline --only-batch"
python bench.py $S --sweep 1,2,3,4,8,16,32,64,128,256,512,1024,2048,4096 --steps 20 --warmup 3 > $O/sweep_c2.json 2> $O/sweep_c2.err
python bench.py $S --batch 1 --rows 10000000 --dim 1024 --metric IP --sweep 1,2,4,8,16 --steps 10 --warmup 3 > $O/small_c3.json 2> $O/small_c3.err
python bench.py $S --batch 1 --rows 12500000 --dim 768 --k 100 --sweep 1,2,4,8,16 --steps 10 --warmup 3 > $O/small_c5.json 2> $O/small_c5.err
python bench.py $S --batch 1 --rows 2500000 --dim 3072 --k 50 --sweep 1,2,4,8,16 --steps 10 --warmup 3 > $O/small_c4.json 2> $O/small_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 1 -c 1 -f -o $O/prof_k3_m128_b1 python bench.py --batch 1 --only-batch --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_b1.log 2>&1
ls -la $O | head -30
