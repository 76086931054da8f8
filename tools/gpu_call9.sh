#!/bin/bash
# 1 GPU: scoring tests, serving latency, ncu full capture of the current pair kernel (user + item launches)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "scoring or recommend or save_load or ids" > gpurun_out/c9_pytest.log 2>&1
tail -n 3 gpurun_out/c9_pytest.log
timeout 600 python - > gpurun_out/c9_latency.json 2> gpurun_out/c9_latency.err <<'PY'
import json, time, numpy as np, sys
sys.path.insert(0, ".")
import pio_b200
from pio_b200 import native, synth
out = {}
for ni in (100_000, 1_000_000):
    itf = np.ascontiguousarray(np.resize(synth.synth_init_factors(1 << 16, 64, 4, 1), (ni, 64)))
    uf = synth.synth_init_factors(1000, 64, 4, 0)
    m = native.NativeALS.from_factors(uf, itf)
    users = np.arange(16, dtype=np.int32)
    q = np.array([5, 77, 4242], np.int32)
    for name, fn in (("recommend_1", lambda: m.recommend(users[:1], 10)), ("recommend_16", lambda: m.recommend(users, 10)),
                     ("similar_3items", lambda: m.similar(q, 20))):
        for _ in range(5): fn()
        ts = []
        for _ in range(50):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        out[f"{name}_{ni}_items_us"] = float(np.median(ts) * 1e6)
    m.close()
print(json.dumps(out))
PY
cat gpurun_out/c9_latency.json
PIO_ALS_TC=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:als_solve_pair -s 6 -c 3 -o gpurun_out/r02_pair_w4_full -f \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-topk --no-parity > gpurun_out/c9_ncu.log 2>&1
tail -2 gpurun_out/c9_ncu.log | cut -c1-300
