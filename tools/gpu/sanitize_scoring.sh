#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over every scoring path on small models
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import pio_b200  # noqa: F401
from pio_b200 import native, synth
rng = np.random.default_rng(1)
for k, ni, nu in ((64, 3000, 80), (20, 1500, 40)):
    itf = synth.synth_init_factors(ni, k, 5, 1); uf = synth.synth_init_factors(nu, k, 6, 0)
    ih = (np.arange(ni) % 9 != 2).astype(np.uint8)
    m = native.NativeALS.from_factors(uf, itf, None, ih)
    w = np.ones(ni); w[::7] = 0.5
    mask = (np.arange(ni) % 5 == 0).astype(np.uint8)
    for topk in (10, 128):
        m.recommend(np.array([3], np.int32), topk, mask, w)                 # fused single query (dot)
        m.similar(np.array([4, 9, 11], np.int32), topk, mask, w)            # fused single query (cosine, 4 vectors)
        m.similar(np.arange(12, dtype=np.int32), topk)                      # three-launch path (> 8 items)
    m.recommend(np.arange(5, dtype=np.int32), 10)                           # 2..16 users
    m.recommend(np.arange(nu, dtype=np.int32), 10, mask, w)                 # blocked batch (dot)
    m.recommend(np.arange(nu, dtype=np.int32), 40)                          # first-generation batch kernel (topk > 32)
    qs = [list(rng.integers(0, ni, rng.integers(0, 7))) for _ in range(37)]
    m.similar_batch(qs, 20, mask, w)                                        # blocked batch (cosine)
    m.similar_batch(qs, 40)                                                 # first-generation batch kernel
    m.close()
print("done")
PY
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/san_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|done" gpurun_out/san_$tool.log | head -12
done
