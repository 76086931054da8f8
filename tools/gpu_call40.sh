#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "batched_recommend" > gpurun_out/c40_pytest.log 2>&1
tail -n 3 gpurun_out/c40_pytest.log
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
cat > /tmp/rb.py <<'PY'
import sys, time, numpy as np, os
sys.path.insert(0, os.getcwd())
import pio_b200
from pio_b200 import native, synth
k, ni, nu = 64, 100_000, 20_000
itf = synth.synth_init_factors(ni, k, 5, 1); uf = synth.synth_init_factors(nu, k, 6, 0)
users = np.arange(nu, dtype=np.int32)
m = native.NativeALS.from_factors(uf, itf, None, None)
m.recommend(users, 10)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_dot_blocked -c 1 -f -o gpurun_out/c40_blocked python /tmp/rb.py > gpurun_out/c40.log 2>&1
