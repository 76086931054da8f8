#!/bin/bash
# 8 GPUs: C2 strong scaling at N = 8 (with the ingest phase trace of the end-to-end call) and the C3 configuration
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 "$@" > gpurun_out/scale8_$name.json 2> gpurun_out/scale8_$name.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale8_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$name", d["value"], d["ms_per_step"], "user", r["ms_per_launch"], "item", r["other_half_step"]["ms_per_launch"], "gram", r["gram_ms_per_iteration"], "comm", r["comm_ms_per_iteration"], d["factor_checksum"], "e2e", d.get("e2e",{}).get("value"), d.get("e2e",{}).get("ingest_ms"), d.get("parity",{}).get("frob_rel"), d.get("parity",{}).get("ok"), "ingest", d["ingest_ms"], "nnz", d["nnz_after_dedup"])
except Exception as e:
    print("$name ERR", e); print(open("gpurun_out/scale8_$name.err").read()[-3000:])
PY
}
PIO_ALS_INGEST_TRACE=1 run c2n8 --steps 10 --warmup 3 --no-topk
grep "ingest r0" gpurun_out/scale8_c2n8.err | tail -34
run c3n8 --workload c3 --steps 3 --warmup 1 --no-topk
