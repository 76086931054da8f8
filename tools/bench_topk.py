"""Serving-side measurement: top-k scoring through the C ABI with host buffers.
python tools/bench_topk.py [n_items] [n_users]  ->  queries/s for recommend (dot) and similar (sum of cosines)."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import pio_b200
from pio_b200 import native
ni = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nu = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
k, nnz, topk = 64, 20 * max(ni, nu), 20
du = torch.empty(nnz, dtype=torch.int32, device="cuda"); di = torch.empty_like(du); dr = torch.empty(nnz, dtype=torch.float32, device="cuda")
native.synth_ratings_device(0, nu, ni, nnz, 3, True, 0, du.data_ptr(), di.data_ptr(), dr.data_ptr())
m = native.NativeALS(k, nu, ni, lam=0.01, implicit=True, init_mode=native.INIT_HASH, seed=3)
m.set_ratings_device(du.data_ptr(), di.data_ptr(), dr.data_ptr(), nnz, dedup=1)
m.run(1)
out = {"n_items": ni, "rank": k, "topk": topk, "item_matrix_mb": ni * k * 4 / 1e6}
rng = np.random.default_rng(0)
for nq in (1, 16, 256):
    users = rng.integers(0, nu, nq).astype(np.int32)
    m.recommend(users, topk)
    torch.cuda.synchronize()
    reps = 20 if nq < 256 else 5
    t0 = time.perf_counter()
    for _ in range(reps):
        m.recommend(users, topk)
    dt = (time.perf_counter() - t0) / reps
    out[f"recommend_batch{nq}"] = {"ms_per_call": dt * 1e3, "queries_per_s": nq / dt,
                                  "item_matrix_gbs": nq * ni * k * 4 / dt / 1e9}
for nqi in (1, 3, 10):
    q = rng.integers(0, ni, nqi).astype(np.int32)
    m.similar(q, topk)
    t0 = time.perf_counter()
    for _ in range(20):
        m.similar(q, topk)
    dt = (time.perf_counter() - t0) / 20
    out[f"similar_{nqi}_query_items"] = {"ms_per_call": dt * 1e3, "queries_per_s": 1 / dt}
print(json.dumps(out))
