#!/bin/bash
# r02 call 11: A/B on ONE box - round-1 library (profiles/r02/variants/libavs_r01.so) vs the current one, same bench harness
O=gpurun_out/c11; mkdir -p $O
S="--no-cpu-baseline --legs none --no-sustained --steps 20 --warmup 5"
for rep in 1 2; do
for v in r01 cur; do
  if [ $v = r01 ]; then export AVS_LIB=$PWD/profiles/r02/variants/libavs_r01.so; else unset AVS_LIB; fi
  timeout 300 python bench.py $S --sweep 1,8,64,128,256,512,2048,4096 > $O/c2_${v}_$rep.json 2> $O/c2_${v}_$rep.err; echo "c2 $v $rep rc=$?"
  timeout 300 python bench.py $S --rows 12500000 --k 100 --sweep 1 > $O/c5_${v}_$rep.json 2> $O/c5_${v}_$rep.err; echo "c5 $v $rep rc=$?"
done
done
for v in r01 cur; do
  if [ $v = r01 ]; then export AVS_LIB=$PWD/profiles/r02/variants/libavs_r01.so; else unset AVS_LIB; fi
  timeout 300 python bench.py $S --rows 10000000 --dim 1024 --metric IP --batch 4096 --sweep 1,16 --steps 10 > $O/c3_${v}.json 2> $O/c3_${v}.err; echo "c3 $v rc=$?"
  timeout 300 python bench.py $S --rows 2500000 --dim 3072 --k 50 --sweep 1 --steps 10 > $O/c4_${v}.json 2> $O/c4_${v}.err; echo "c4 $v rc=$?"
  timeout 300 python bench.py $S --rows 125000 --sweep 1,8 > $O/shard125k_${v}.json 2> $O/shard125k_${v}.err; echo "125k $v rc=$?"
done
unset AVS_LIB
timeout 300 python profiles/r02/trace_levels.py --batches 1024 --rows 12500000 --k 100 > $O/trace_c5_b1024.json 2> $O/trace.err; echo "trace c5 rc=$?"
