#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/c4_pytest.log 2>&1
tail -n 4 gpurun_out/c4_pytest.log
for cfg in "c2pair:c2:PIO_ALS_TC=0" "c2default:c2:" "small128:small128:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; wl=${rest%%:*}; envs=${rest#*:}
  env $envs timeout 900 python bench.py --workload $wl --steps 5 --warmup 2 --no-e2e --no-cpu-baseline --no-topk > gpurun_out/c4_bench_$name.json 2> gpurun_out/c4_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c4_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["other_half_step"]["ms_per_launch"], d["factor_checksum"], d.get("parity",{}).get("frob_rel"), d.get("parity",{}).get("ok"))
except Exception as e:
    print("$name ERR", e); print(open("gpurun_out/c4_bench_$name.err").read()[-1500:])
PY
done
