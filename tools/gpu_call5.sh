#!/bin/bash
# 2 GPUs: full GPU test suite (incl. sharded == single GPU), C2 bench at N = 1 and N = 2
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/c5_pytest.log 2>&1
tail -n 5 gpurun_out/c5_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-topk > gpurun_out/c5_bench_n1.json 2> gpurun_out/c5_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-topk > gpurun_out/c5_bench_n2.json 2> gpurun_out/c5_bench_n2.err
for n in n1 n2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c5_bench_$n.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$n", d["value"], d["ms_per_step"], "user", r["ms_per_launch"], "item", r["other_half_step"]["ms_per_launch"], "gram", r["gram_ms_per_iteration"], "comm", r["comm_ms_per_iteration"], d["factor_checksum"], "e2e", d.get("e2e",{}).get("value"), d.get("e2e",{}).get("ingest_ms"), d.get("parity",{}).get("frob_rel"), d.get("parity",{}).get("ok"), "ingest", d["ingest_ms"])
except Exception as e:
    print("$n ERR", e); print(open("gpurun_out/c5_bench_$n.err").read()[-2500:])
PY
done
