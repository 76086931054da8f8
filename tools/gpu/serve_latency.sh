#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_templates.py -q -m gpu -x -k "single_query or scoring_weights" > gpurun_out/serve_pytest.log 2>&1
tail -n 3 gpurun_out/serve_pytest.log
timeout 300 python tools/serve_latency.py > gpurun_out/serve_lat_1m.json 2> gpurun_out/serve_lat.err; cat gpurun_out/serve_lat_1m.json
timeout 300 python tools/serve_latency.py --items 100000 > gpurun_out/serve_lat_100k.json 2>> gpurun_out/serve_lat.err; cat gpurun_out/serve_lat_100k.json
PIO_ALS_SERVE_TRACE=1 timeout 300 python tools/serve_latency.py --calls 3 > /dev/null 2> gpurun_out/serve_trace_1m.err; grep "serve trace" gpurun_out/serve_trace_1m.err | sed -n '21,23p;44,45p'
PIO_ALS_SERVE_TRACE=1 timeout 300 python tools/serve_latency.py --calls 3 --items 100000 > /dev/null 2> gpurun_out/serve_trace_100k.err; grep "serve trace" gpurun_out/serve_trace_100k.err | sed -n '21,23p;44,45p'
