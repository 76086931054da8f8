#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import torch, time
x=torch.empty(300_000_000, dtype=torch.int32).pin_memory()
d=torch.empty_like(x, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); print("pinned H2D GB/s", x.numel()*4/(time.perf_counter()-t)/1e9)
y=torch.empty(300_000_000, dtype=torch.int32)
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(y); torch.cuda.synchronize(); print("pageable H2D GB/s", y.numel()*4/(time.perf_counter()-t)/1e9)
PY
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
PIO_ALS_INGEST_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-topk --no-parity --no-cpu-baseline > gpurun_out/c33_trace_n1.json 2> gpurun_out/c33_trace_n1.err
grep "ingest r0" gpurun_out/c33_trace_n1.err | tail -22
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c33_trace_n1.json").read().strip().splitlines()[-1])
print(d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ingest_ms"], d["e2e"]["run_ms"], d["e2e"]["seconds_per_train_call"])
PY
