# tuning experiments on the mid-batch regime (one store per process, options via --opt); output: one JSON line per variant
O=gpurun_out/r01c; mkdir -p $O
C2="--steps 20 --warmup 3 --no-cpu-baseline --only-batch --batch 128"
C3="--rows 10000000 --dim 1024 --metric IP --steps 10 --warmup 3 --no-cpu-baseline --only-batch --batch 128"
run() { name=$1; shift; python bench.py "$@" > $O/$name.json 2> $O/$name.err; }
run c2_base   $C2 --sweep 2,3,4,16,64,96,128,192
run c2_cg1    $C2 --sweep 4,16,64,96,128 --opt cta_group_small=1
run c2_sig3   $C2 --sweep 2,3,4,16,64,96,128,192 --opt coarse_sigma=3
run c2_fine64 $C2 --sweep 64,96,128,192 --opt fine_min_batch=64
run c2_cg1s3  $C2 --sweep 4,16,64,96,128 --opt cta_group_small=1 --opt coarse_sigma=3
run c2_gemm2  $C2 --sweep 2,3 --opt gemm_min_batch=2
run c2_gemm2c $C2 --sweep 2,3 --opt gemm_min_batch=2 --opt cta_group_small=1 --opt coarse_sigma=3
run c3_base   $C3 --sweep 4,16,64,128
run c3_cg1    $C3 --sweep 4,16,64,128 --opt cta_group_small=1
run c3_cg1s3  $C3 --sweep 4,16,64,128 --opt cta_group_small=1 --opt coarse_sigma=3
run c3_fine64 $C3 --sweep 64,128 --opt fine_min_batch=64
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
