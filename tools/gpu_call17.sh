#!/bin/bash
# 1 GPU: fused single-query kernel v2 (bulk-copy ring, compile-time row width, pruned list merge)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "single_query or recommend_and_similar or scoring_weights" > gpurun_out/c17_pytest.log 2>&1
tail -n 5 gpurun_out/c17_pytest.log
timeout 300 python tools/serve_latency.py > gpurun_out/c17_lat_1m.json 2> gpurun_out/c17_lat.err; cat gpurun_out/c17_lat_1m.json
timeout 300 python tools/serve_latency.py --items 100000 > gpurun_out/c17_lat_100k.json 2>> gpurun_out/c17_lat.err; cat gpurun_out/c17_lat_100k.json
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max,smsp__inst_executed.sum --clock-control none -k regex:score_one -c 60 --csv --log-file gpurun_out/c17_launches.csv python tools/serve_latency.py --calls 5 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open("gpurun_out/c17_launches.csv") if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); mi=h.index("Metric Name"); ii=h.index("ID")
seen={}
for r in rows[1:]:
    seen.setdefault(r[ii], [r[ki][:34]]).append(r[vi])
for k in list(seen)[::6][:10]: print(seen[k])
PY
