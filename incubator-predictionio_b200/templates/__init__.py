"""Engine templates re-expressed over pio_b200.mllib (the native-als module): same DASE classes,
Params, queries and results as the reference's examples/scala-parallel-* templates; only the MLlib
call inside `train` (and the per-query scans inside `predict`) go to the GPU."""
