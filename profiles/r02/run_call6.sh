#!/bin/bash
O=gpurun_out/c6; mkdir -p $O
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 128,1024 > $O/trace_epi_c2.json 2> $O/trace.err; echo "rc=$?"
timeout 300 python profiles/r02/trace_levels.py --per-cta --batches 1024 --opt final_sigma=8 > $O/trace_epi_c2_sigma8.json 2>> $O/trace.err; echo "rc=$?"
tail -n 3 $O/trace.err
