#!/bin/bash
# first GPU call of round 2: peaks, lockstep solver, full GPU test suite, A/B of the rank-64 kernels at C2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 120 tools/peaks > gpurun_out/c1_peaks.json 2> gpurun_out/c1_peaks.err
timeout 600 python -m pytest tests/test_gpu_lockstep.py -x -q -m gpu > gpurun_out/c1_lockstep.log 2>&1
timeout 300 python tools/bench_solver.py > gpurun_out/c1_solver.json 2> gpurun_out/c1_solver.err
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c1_pytest.log 2>&1
for cfg in "default:" "pair:PIO_ALS_TC=0" "oldmma:PIO_ALS_TC=0 PIO_ALS_MMA=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu-baseline --no-topk > gpurun_out/c1_bench_$name.json 2> gpurun_out/c1_bench_$name.err
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench_full.json 2> gpurun_out/c1_bench_full.err
echo finished > gpurun_out/c1_done.txt
tail -3 gpurun_out/c1_lockstep.log gpurun_out/c1_pytest.log
cat gpurun_out/c1_peaks.json gpurun_out/c1_solver.json
for f in gpurun_out/c1_bench_*.json; do echo $f; python - <<PY
import json,sys
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","factor_checksum")}, d["roofline"]["ms_per_launch"], d["roofline"]["other_half_step"]["ms_per_launch"], d.get("parity",{}).get("frob_rel"), d.get("parity",{}).get("ok"))
except Exception as e:
    print("ERR", e)
PY
done
