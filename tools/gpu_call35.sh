#!/bin/bash
mkdir -p gpurun_out
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_dot_blocked -c 1 -f -o gpurun_out/c35_blocked python /tmp/rb.py > gpurun_out/c35.log 2>&1
ls -la gpurun_out/*.ncu-rep
