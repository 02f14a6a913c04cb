# marginal cost of an accepted row in the tensor-core scan (final_sigma sweep) + source-level ncu capture of the final level
O=gpurun_out/r01k; mkdir -p $O
B="--steps 30 --warmup 5 --no-cpu-baseline --only-batch --batch 1024"
run() { name=$1; shift; timeout 300 python bench.py "$@" > $O/$name.json 2> $O/$name.err; }
run s2  $B
run s4  $B --opt final_sigma=4
run s8  $B --opt final_sigma=8
run s16 $B --opt final_sigma=16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 14 -c 1 -f -o $O/prof_k3_b1024_src python bench.py --batch 1024 --only-batch --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu.log 2>&1
ls -la $O
