#!/bin/bash
# 1 GPU: full test suite, full default bench (the driver's command), reference arm, ncu launch list of the C2 step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/c32_pytest.log 2>&1
tail -n 4 gpurun_out/c32_pytest.log
timeout 1200 python bench.py > gpurun_out/c32_bench.json 2> gpurun_out/c32_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c32_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(d["value"], d["ms_per_step"], "user", r["ms_per_launch"], "item", r["other_half_step"]["ms_per_launch"], "gram", r["gram_ms_per_iteration"], d["factor_checksum"], "e2e", d["e2e"]["value"], d["e2e"].get("ingest_ms"), "parity", d["parity"]["frob_rel"], d["parity"]["ok"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
t=d["topk"]
print("recommend", t["recommend"]["value"], t["recommend"]["single_query_ms"], t["recommend"]["bit_exact_vs_oracle_on_sample"])
print("similar", t["similar_c4"].get("value"), t["similar_c4"].get("single_query_ms"), t["similar_c4"].get("bit_exact_vs_oracle_on_sample"), t["similar_c4"].get("error"))
print("nb", t["naive_bayes_c5"])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c32_ref.json 2> gpurun_out/c32_ref.err; cat gpurun_out/c32_ref.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c32_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-topk --no-parity > gpurun_out/c32_ncu.log 2>&1
tail -n 2 gpurun_out/c32_ncu.log | cut -c1-300
