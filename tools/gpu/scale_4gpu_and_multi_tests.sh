#!/bin/bash
# 4 GPUs: C2 at N = 4 (the missing point of the scaling table) + the multi-GPU tests (world size 2)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 10 --warmup 3 --no-topk > gpurun_out/scale4_c2n4.json 2> gpurun_out/scale4_c2n4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/scale4_c2n4.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(d["value"], d["ms_per_step"], "user", r["ms_per_launch"], "item", r["other_half_step"]["ms_per_launch"], "gram", r["gram_ms_per_iteration"], "comm", r["comm_ms_per_iteration"], d["factor_checksum"], "e2e", d["e2e"]["value"], d["e2e"]["ingest_ms"], d["parity"]["frob_rel"], d["clocks"])
PY
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/scale4_multi.log 2>&1
tail -n 3 gpurun_out/scale4_multi.log
