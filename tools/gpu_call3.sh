#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/peaks > gpurun_out/c3_peaks.json 2> gpurun_out/c3_peaks.err
export PIO_ALS_TC=0
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "heavy or rank64 or tolerance or one_iteration or fixtures or negative" > gpurun_out/c3_pytest.log 2>&1
for w in 1 2 4 6 12; do
  PIO_ALS_PAIR_WARPS=$w timeout 600 python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu-baseline --no-topk --no-parity > gpurun_out/c3_bench_w$w.json 2> gpurun_out/c3_bench_w$w.err
done
cat gpurun_out/c3_peaks.json; tail -n 3 gpurun_out/c3_pytest.log
for w in 1 2 4 6 12; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c3_bench_w$w.json").read().strip().splitlines()[-1])
    print("w$w", d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["other_half_step"]["ms_per_launch"], d["factor_checksum"])
except Exception as e:
    print("ERR", e)
PY
done
