#!/bin/bash
# 2 GPUs: YtY by rank-local classes -- bit-identity across world sizes, parity, timing at N = 1 / 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/c44_multi.log 2>&1
tail -n 3 gpurun_out/c44_multi.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "one_iteration or trained_factors or deterministic or heavy_rows or fixtures" > gpurun_out/c44_parity.log 2>&1
tail -n 3 gpurun_out/c44_parity.log
for n in 1 2; do
if [ $n = 1 ]; then timeout 600 python bench.py --steps 10 --warmup 3 --no-topk --no-e2e --no-cpu-baseline > gpurun_out/c44_n$n.json 2> gpurun_out/c44_n$n.err
else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $n --steps 10 --warmup 3 --no-topk --no-e2e > gpurun_out/c44_n$n.json 2> gpurun_out/c44_n$n.err; fi
python - <<PY
import json
d=json.loads(open("gpurun_out/c44_n$n.json").read().strip().splitlines()[-1])
r=d["roofline"]
print($n, d["value"], d["ms_per_step"], "user", r["ms_per_launch"], "item", r["other_half_step"]["ms_per_launch"], "gram", r["gram_ms_per_iteration"], "comm", r["comm_ms_per_iteration"], d["factor_checksum"], d["parity"]["frob_rel"], d["parity"]["ok"])
PY
done
