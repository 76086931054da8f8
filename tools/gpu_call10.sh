#!/bin/bash
mkdir -p gpurun_out
PIO_ALS_TC=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:als_solve_pair -s 3 -c 3 -o gpurun_out/r02_pair_w4_full -f \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-topk --no-parity > gpurun_out/c10_ncu.log 2>&1
tail -2 gpurun_out/c10_ncu.log | cut -c1-200
ls -la gpurun_out/r02_pair_w4_full.ncu-rep
