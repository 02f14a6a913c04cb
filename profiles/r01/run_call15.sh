# A/B: same shared-memory carve-out for every kernel of the pipeline (AVS_CARVEOUT=1, default) vs the driver's per-kernel choice (0)
O=gpurun_out/r01p; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
S="--no-cpu-baseline --only-batch --steps 30 --warmup 5"
for rep in 1 2; do for co in 0 1; do
  AVS_CARVEOUT=$co python bench.py $S --sweep 1,8,64,1024 > $O/c2_co${co}_$rep.json 2> $O/c2_co${co}_$rep.err
  AVS_CARVEOUT=$co python bench.py $S --rows 125000 --sweep 1,1024 > $O/sh_co${co}_$rep.json 2> $O/sh_co${co}_$rep.err
done; done
