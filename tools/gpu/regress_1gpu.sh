#!/bin/bash
# 1 GPU: full test suite + full default bench after the scoring-kernel work
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/regress_pytest.log 2>&1
tail -n 4 gpurun_out/regress_pytest.log
timeout 1200 python bench.py > gpurun_out/regress_bench.json 2> gpurun_out/regress_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/regress_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(d["value"], d["ms_per_step"], "user", r["ms_per_launch"], "item", r["other_half_step"]["ms_per_launch"], "gram", r["gram_ms_per_iteration"], d["factor_checksum"], "e2e", d["e2e"]["value"], d["e2e"].get("ingest_ms"), "parity", d["parity"]["frob_rel"], d["parity"]["ok"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
t=d["topk"]
print("recommend", t["recommend"]["value"], t["recommend"]["single_query_ms"], t["recommend"]["bit_exact_vs_oracle_on_sample"], t["recommend"]["roofline"]["frac"])
print("similar", t["similar_c4"].get("value"), t["similar_c4"].get("single_query_ms"), t["similar_c4"].get("bit_exact_vs_oracle_on_sample"), t["similar_c4"].get("roofline",{}).get("frac"), t["similar_c4"].get("error"))
PY
python -c "
import sys; sys.path.insert(0,'.')
import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
