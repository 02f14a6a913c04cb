#!/bin/bash
O=gpurun_out/c5; mkdir -p $O
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 128,1024 > $O/trace_percta_c2.json 2> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1024 --rows 125000 > $O/trace_percta_125k.json 2>> $O/trace.err; echo "rc=$?"
tail -n 3 $O/trace.err
