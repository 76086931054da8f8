#!/bin/bash
# 8 GPUs: C2 at N = 8 with the rank-local YtY classes (overlapped with the factor exchange)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 --no-topk > gpurun_out/c45_c2n8.json 2> gpurun_out/c45_c2n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c45_c2n8.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(d["value"], d["ms_per_step"], "user", r["ms_per_launch"], "item", r["other_half_step"]["ms_per_launch"], "gram", r["gram_ms_per_iteration"], "comm", r["comm_ms_per_iteration"], d["factor_checksum"], "e2e", d["e2e"]["value"], d["e2e"]["ingest_ms"], d["parity"]["frob_rel"], d["clocks"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 10 --warmup 3 --no-topk --no-e2e --no-parity > gpurun_out/c45_c2n2.json 2> gpurun_out/c45_c2n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c45_c2n2.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("N=2", d["value"], d["ms_per_step"], "user", r["ms_per_launch"], "item", r["other_half_step"]["ms_per_launch"], "gram", r["gram_ms_per_iteration"], "comm", r["comm_ms_per_iteration"], d["factor_checksum"])
PY
