#!/bin/bash
# The ncu captures behind profiles/r02_*: one kernel per invocation (a report is 30-45 MB; gpurun returns at most 64 MiB).
#   tools/gpu/profile_kernels.sh pair | launches | one | blocked
# Run under gpurun on ONE GPU; read the report here with `ncu -i gpurun_out/<name>.ncu-rep --page raw --csv`
# (or --page source --csv --print-source sass).  Numbers printed by a run under ncu are never bench values.
mkdir -p gpurun_out
case "$1" in
  pair)      # second iteration's item parts, item whole rows, user rows of the C2 step
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:als_solve_pair -s 3 -c 3 -f -o gpurun_out/pair_full \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-topk --no-parity ;;
  launches)  # launch list of the same command (profiles/r02_launches_c2.csv)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-topk --no-parity ;;
  one)       # the fused single-query kernel: last dot-product query and first cosine query of tools/serve_latency.py
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_one -s 24 -c 2 -f -o gpurun_out/score_one \
      python tools/serve_latency.py --calls 5 ;;
  blocked)   # the blocked batched recommend kernel, 20 k users x 100 k items
    cat > /tmp/rb.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import pio_b200  # noqa: F401
from pio_b200 import native, synth
k, ni, nu = 64, 100_000, 20_000
m = native.NativeALS.from_factors(synth.synth_init_factors(nu, k, 6, 0), synth.synth_init_factors(ni, k, 5, 1), None, None)
m.recommend(np.arange(nu, dtype=np.int32), 10)
PY
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_dot_blocked -c 1 -f -o gpurun_out/dot_blocked \
      python /tmp/rb.py ;;
  *) echo "usage: $0 pair|launches|one|blocked"; exit 2 ;;
esac
ls -la gpurun_out/
