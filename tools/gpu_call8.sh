#!/bin/bash
# 1 GPU: full test suite, full bench (topk + C4 + C5), small128 launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/c8_pytest.log 2>&1
tail -n 4 gpurun_out/c8_pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c8_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print(d["value"], d["ms_per_step"], "user", r["ms_per_launch"], "item", r["other_half_step"]["ms_per_launch"], "gram", r["gram_ms_per_iteration"], "e2e", d["e2e"]["value"], d["parity"]["frob_rel"], d["parity"]["ok"])
    print(json.dumps(d["topk"], indent=1)[:6000])
    print(d["cpu_baseline"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/c8_bench.err").read()[-3000:])
PY
