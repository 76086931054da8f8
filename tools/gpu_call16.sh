#!/bin/bash
# 1 GPU: full ncu capture of the fused single-query kernel at 1 M and 100 k items
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_one -s 25 -c 2 -f -o gpurun_out/c16_one_1m python tools/serve_latency.py --calls 5 > gpurun_out/c16_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_one -s 25 -c 2 -f -o gpurun_out/c16_one_100k python tools/serve_latency.py --calls 5 --items 100000 > gpurun_out/c16_b.log 2>&1
ls -la gpurun_out/c16*
