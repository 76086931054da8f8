#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_one -s 24 -c 2 -f -o gpurun_out/c20_one_b1 python tools/serve_latency.py --calls 5 --bulk 1 > gpurun_out/c20_b.log 2>&1
ls -la gpurun_out/
