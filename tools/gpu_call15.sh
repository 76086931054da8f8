#!/bin/bash
# 1 GPU: fused single-query kernel -- tests, latency (fused / three launches), kernel durations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "single_query or recommend_and_similar or scoring_weights" > gpurun_out/c15_pytest.log 2>&1
tail -n 5 gpurun_out/c15_pytest.log
for f in 1 0; do
  timeout 300 python tools/serve_latency.py --fused $f > gpurun_out/c15_lat_1m_f$f.json 2> gpurun_out/c15_lat.err; cat gpurun_out/c15_lat_1m_f$f.json
  timeout 300 python tools/serve_latency.py --fused $f --items 100000 > gpurun_out/c15_lat_100k_f$f.json 2>> gpurun_out/c15_lat.err; cat gpurun_out/c15_lat_100k_f$f.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:score_one -c 60 --csv --log-file gpurun_out/c15_launches.csv python tools/serve_latency.py --calls 5 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open("gpurun_out/c15_launches.csv") if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[1::5][:12]: print(r[ki][:50], r[vi])
PY
