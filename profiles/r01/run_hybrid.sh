# hybrid small-batch pipeline (<= 8 queries: dense warp-dot level, tensor-core later levels) against the previous auto mode
O=gpurun_out/r01h; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
B="--steps 30 --warmup 5 --no-cpu-baseline --only-batch --batch 1"
run() { name=$1; shift; timeout 300 python bench.py "$@" > $O/$name.json 2> $O/$name.err; }
run c2_hyb   $B --sweep 1,2,3,4,8,16
run c2_old   $B --sweep 1,2,3,4,8 --opt hybrid=0
run c3_hyb   $B --steps 10 --rows 10000000 --dim 1024 --metric IP --sweep 1,2,4,8
run c5_hyb   $B --steps 10 --rows 12500000 --dim 768 --k 100 --sweep 1,2,8
run c5_old   $B --steps 10 --rows 12500000 --dim 768 --k 100 --sweep 1,2,8 --opt hybrid=0
run sh_hyb   $B --rows 125000 --sweep 1,2,4,8
run sh_old   $B --rows 125000 --sweep 1,2,4,8 --opt hybrid=0
run c4_hyb   $B --steps 10 --rows 2500000 --dim 3072 --k 50 --sweep 1,2,8
