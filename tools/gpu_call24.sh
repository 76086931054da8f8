#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "single_query or recommend_and_similar or scoring_weights" > gpurun_out/c24_pytest.log 2>&1
tail -n 5 gpurun_out/c24_pytest.log
timeout 300 python tools/serve_latency.py > gpurun_out/c24_lat_1m.json 2> gpurun_out/c24_lat.err; cat gpurun_out/c24_lat_1m.json
timeout 300 python tools/serve_latency.py --items 100000 > gpurun_out/c24_lat_100k.json 2>> gpurun_out/c24_lat.err; cat gpurun_out/c24_lat_100k.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_one -s 24 -c 2 -f -o gpurun_out/c24_one python tools/serve_latency.py --calls 5 > gpurun_out/c24_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_one -s 24 -c 1 -f -o gpurun_out/c24_one_100k python tools/serve_latency.py --calls 5 --items 100000 > gpurun_out/c24_c.log 2>&1
