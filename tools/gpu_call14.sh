#!/bin/bash
# 1 GPU: rank-128 finish kernel with six synchronised warps per CTA -- tests + small128 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lockstep.py tests/test_gpu_parity.py -q -m gpu -x -k "128 or lockstep or rank" > gpurun_out/c14_pytest.log 2>&1
tail -n 3 gpurun_out/c14_pytest.log
timeout 600 python bench.py --workload small128 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-topk > gpurun_out/c14_small128.json 2> gpurun_out/c14_small128.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c14_small128.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("parity",{}).get("frob_rel"), d["roofline"])
PY
