#!/bin/bash
# 2 GPUs: ingest phase trace (fine marks) at N = 2, plus the multi-GPU test
mkdir -p gpurun_out
PIO_ALS_INGEST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 2 --no-topk --no-parity > gpurun_out/c13_trace_n2.json 2> gpurun_out/c13_trace_n2.err
grep "ingest r0" gpurun_out/c13_trace_n2.err | tail -32
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c13_trace_n2.json").read().strip().splitlines()[-1])
r=d["roofline"]; print(d["value"], d["ms_per_step"], "gram", r["gram_ms_per_iteration"], "comm", r["comm_ms_per_iteration"], "e2e", d["e2e"]["value"], d["e2e"]["ingest_ms"], d["factor_checksum"])
PY
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/c13_multi.log 2>&1
tail -n 3 gpurun_out/c13_multi.log
