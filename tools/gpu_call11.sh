#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "heavy or rank64 or tolerance or one_iteration or fixtures or negative or ragged or single" > gpurun_out/c11_pytest.log 2>&1
tail -n 3 gpurun_out/c11_pytest.log
for cfg in "duo:" "pair:PIO_ALS_DUO=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu-baseline --no-topk > gpurun_out/c11_bench_$name.json 2> gpurun_out/c11_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c11_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["other_half_step"]["ms_per_launch"], d["factor_checksum"], d.get("parity",{}).get("frob_rel"), d.get("parity",{}).get("ok"))
except Exception as e:
    print("$name ERR", e); print(open("gpurun_out/c11_bench_$name.err").read()[-1500:])
PY
done
