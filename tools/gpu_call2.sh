#!/bin/bash
# ncu captures of the pair kernel at C2 (both sides on the pair kernel)
mkdir -p gpurun_out
export PIO_ALS_TC=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:als_solve_pair -s 4 -c 2 -o gpurun_out/r02_pair_full -f \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-topk --no-parity > gpurun_out/c2_ncu.log 2>&1
tail -5 gpurun_out/c2_ncu.log
ls -la gpurun_out/*.ncu-rep
