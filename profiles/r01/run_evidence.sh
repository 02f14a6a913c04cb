set -x
O=gpurun_out/r01b; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --sweep 1,2,3,4,8,16,32,64,128,256,512,1024,2048,4096 --steps 20 --warmup 3 --no-cpu-baseline > $O/sweep_c2.json 2> $O/sweep_c2.err
python bench.py --rows 10000000 --dim 1024 --metric IP --sweep 1,2,4,8,16,32,64,128,256,512,1024,2048,4096 --steps 10 --warmup 3 --no-cpu-baseline > $O/sweep_c3.json 2> $O/sweep_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
tail -3 $O/pytest_gpu.log; cat $O/bench_default.json | cut -c1-600
