O=gpurun_out/r01j; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err; cut -c1-200 $O/bench_default.json
B="--steps 30 --warmup 5 --no-cpu-baseline --only-batch --batch 1"
run() { name=$1; shift; timeout 300 python bench.py "$@" > $O/$name.json 2> $O/$name.err; }
run d64k  $B --sweep 1,2,4,8
run d32k  $B --sweep 1,2,4,8 --opt dense_rows=32768
run d48k  $B --sweep 1,2,4,8 --opt dense_rows=49152
run c5_d64k $B --steps 10 --rows 12500000 --dim 768 --k 100 --sweep 1,2
run c5_d32k $B --steps 10 --rows 12500000 --dim 768 --k 100 --sweep 1,2 --opt dense_rows=32768
