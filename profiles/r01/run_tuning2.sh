# tuning experiments, round 2: new small-batch defaults, fine-level ratio, small (1/8 C2) shard breakdown
O=gpurun_out/r01d; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
B="--steps 20 --warmup 3 --no-cpu-baseline --only-batch"
C3="--rows 10000000 --dim 1024 --metric IP"
run() { name=$1; shift; python bench.py "$@" > $O/$name.json 2> $O/$name.err; }
run c2_new      $B --batch 128 --sweep 2,3,4,16,64,128,129,192,256
run c2_fr4      $B --batch 1024 --sweep 256,512,1024,4096
run c2_fr8      $B --batch 1024 --sweep 256,512,1024,4096 --opt fine_ratio=8
run c2_fr8s3    $B --batch 1024 --sweep 256,512,1024,4096 --opt fine_ratio=8 --opt final_sigma=3
run sh_fr4      $B --rows 125000 --batch 1024 --sweep 1,256,1024
run sh_fr8      $B --rows 125000 --batch 1024 --sweep 256,1024 --opt fine_ratio=8
run sh_fr16     $B --rows 125000 --batch 1024 --sweep 256,1024 --opt fine_ratio=16
run c3_fr4      $B $C3 --steps 10 --batch 1024 --sweep 256,1024
run c3_fr8      $B $C3 --steps 10 --batch 1024 --sweep 256,1024 --opt fine_ratio=8
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_shard125k.csv python bench.py --rows 125000 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_shard.log 2>&1
