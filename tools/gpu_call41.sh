#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "batched_similar or scoring_weights or batched_recommend" > gpurun_out/c41_pytest.log 2>&1
tail -n 8 gpurun_out/c41_pytest.log
python - <<'PY'
import sys, time, numpy as np, os
sys.path.insert(0, os.getcwd())
import pio_b200
from pio_b200 import native, synth
n_it, kk, nqs = 1_000_000, 64, 10_000
rng = np.random.default_rng(3)
big = np.ascontiguousarray(np.resize(synth.synth_init_factors(1 << 16, kk, 4, 1), (n_it, kk)))
big *= (1.0 + (np.arange(n_it, dtype=np.float32) % 97)[:, None] / 97.0)
queries = [rng.integers(0, n_it, rng.integers(1, 6)).astype(np.int32) for _ in range(nqs)]
for blocked in ("1", "0"):
    os.environ["PIO_ALS_SCORE_BLOCKED"] = blocked
    mm = native.NativeALS.from_factors(None, big, None, None)
    mm.similar_batch(queries[:64], 20)
    t0 = time.perf_counter(); mm.similar_batch(queries, 20); dt = time.perf_counter() - t0
    print("blocked", blocked, "C4 similar_batch:", nqs / dt, "q/s", dt)
    mm.close()
PY
