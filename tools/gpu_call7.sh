#!/bin/bash
# 2 GPUs: tests, ingest phase trace at N = 1 / 2, per-kernel times of the rank-128 path
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/c7_pytest.log 2>&1
tail -n 3 gpurun_out/c7_pytest.log
PIO_ALS_INGEST_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-topk --no-parity > gpurun_out/c7_trace_n1.json 2> gpurun_out/c7_trace_n1.err
PIO_ALS_INGEST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 1 --no-topk --no-parity > gpurun_out/c7_trace_n2.json 2> gpurun_out/c7_trace_n2.err
grep "ingest r0" gpurun_out/c7_trace_n1.err | tail -12
grep "ingest r0" gpurun_out/c7_trace_n2.err | tail -14
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c7_launches_small128.csv python bench.py --workload small128 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-topk --no-parity > gpurun_out/c7_small128.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/c7_launches_small128.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: agg[r[ki][:90]][0]+=1; agg[r[ki][:90]][1]+=float(r[vi].replace(",",""))
    except Exception: pass
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]: print(f"{t/1e6:9.3f} ms {n:4d}  {k}")
PY
