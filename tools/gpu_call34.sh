#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks tools/peaks.cu 2>/dev/null; ./tools/peaks > gpurun_out/c34_peaks.json; cat gpurun_out/c34_peaks.json
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "batched_recommend or scoring_weights or recommend_and_similar" > gpurun_out/c34_pytest.log 2>&1
tail -n 5 gpurun_out/c34_pytest.log
python - <<'PY'
import sys, time, numpy as np, os
sys.path.insert(0, os.getcwd())
import pio_b200
from pio_b200 import native, synth
k, ni, nu = 64, 100_000, 100_000
itf = synth.synth_init_factors(ni, k, 5, 1); uf = synth.synth_init_factors(nu, k, 6, 0)
users = np.arange(nu, dtype=np.int32)
for blocked in ("1", "0"):
    os.environ["PIO_ALS_SCORE_BLOCKED"] = blocked
    m = native.NativeALS.from_factors(uf, itf, None, None)
    m.recommend(users[:1000], 10)
    t0 = time.perf_counter(); m.recommend(users, 10); dt = time.perf_counter() - t0
    print("blocked", blocked, "recommend 100k x top-10 over 100k items:", nu / dt, "pred/s", dt)
    m.close()
PY
