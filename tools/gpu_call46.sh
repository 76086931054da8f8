#!/bin/bash
# 4 GPUs: C2 at N = 4 with the rank-local YtY classes (checksum must equal the 1 / 2 / 8-GPU one)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 10 --warmup 3 --no-topk --no-parity > gpurun_out/c46_c2n4.json 2> gpurun_out/c46_c2n4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c46_c2n4.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(d["value"], d["ms_per_step"], "user", r["ms_per_launch"], "item", r["other_half_step"]["ms_per_launch"], "gram", r["gram_ms_per_iteration"], "comm", r["comm_ms_per_iteration"], d["factor_checksum"], "e2e", d["e2e"]["value"], d["e2e"]["ingest_ms"])
PY
