#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_one -s 24 -c 1 -f -o gpurun_out/c21_one_b0 python tools/serve_latency.py --calls 5 --bulk 0 > gpurun_out/c21_b.log 2>&1
ls -la gpurun_out/
