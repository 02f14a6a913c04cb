# verification of the dense-buffer select fix + small-shard / large-k latency + one ncu capture of the M=128 scan
O=gpurun_out/r01e; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
B="--steps 20 --warmup 3 --no-cpu-baseline --only-batch"
run() { name=$1; shift; timeout 300 python bench.py "$@" > $O/$name.json 2> $O/$name.err; }
run sh_b1     $B --rows 125000 --batch 1 --sweep 1,2,3
run c2_k100   $B --k 100 --batch 1 --sweep 1,2,64,1024
run sh_k100   $B --rows 125000 --k 100 --batch 1 --sweep 1,1024
run small50k  $B --rows 50000 --batch 1 --sweep 1,2,64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 5 -c 1 -f -o $O/prof_k3_cg1_b64 python bench.py --batch 64 --only-batch --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_cg1.log 2>&1
ls -la $O
