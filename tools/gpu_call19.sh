#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "single_query or recommend_and_similar or scoring_weights" > gpurun_out/c19_pytest.log 2>&1
tail -n 5 gpurun_out/c19_pytest.log
PIO_ALS_SERVE_BULK=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "single_query" > gpurun_out/c19_pytest_b0.log 2>&1
tail -n 2 gpurun_out/c19_pytest_b0.log
for b in 1 0; do
timeout 300 python tools/serve_latency.py --bulk $b > gpurun_out/c19_lat_1m_b$b.json 2> gpurun_out/c19_lat.err; cat gpurun_out/c19_lat_1m_b$b.json
timeout 300 python tools/serve_latency.py --bulk $b --items 100000 > gpurun_out/c19_lat_100k_b$b.json 2>> gpurun_out/c19_lat.err; cat gpurun_out/c19_lat_100k_b$b.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_one -s 24 -c 2 -f -o gpurun_out/c19_one_b0 python tools/serve_latency.py --calls 5 --bulk 0 > gpurun_out/c19_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_one -s 24 -c 2 -f -o gpurun_out/c19_one_b1 python tools/serve_latency.py --calls 5 --bulk 1 > gpurun_out/c19_b.log 2>&1
