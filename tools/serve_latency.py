#!/usr/bin/env python
"""Single-query serving latency of the C-ABI scoring calls (pio_als_recommend / pio_als_similar with one query):
median / p10 / p90 over repeated calls on an imported random model.  `--fused 0` measures the three-launch path."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=1_000_000)
    ap.add_argument("--rank", type=int, default=64)
    ap.add_argument("--calls", type=int, default=300)
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--bulk", type=int, default=1)
    args = ap.parse_args()
    os.environ["PIO_ALS_SERVE_FUSED"] = str(args.fused)
    os.environ["PIO_ALS_SERVE_BULK"] = str(args.bulk)
    import pio_b200  # noqa: F401
    from pio_b200 import native, synth
    ni, k = args.items, args.rank
    itf = np.ascontiguousarray(np.resize(synth.synth_init_factors(1 << 16, k, 7, 1), (ni, k)))
    itf *= (1.0 + (np.arange(ni, dtype=np.float32) % 97)[:, None] / 97.0)
    uf = synth.synth_init_factors(1000, k, 8, 0)
    m = native.NativeALS.from_factors(uf, itf, None, None)
    rng = np.random.default_rng(3)
    out = {"items": ni, "rank": k, "fused": args.fused, "calls": args.calls}

    def timed(fn):
        for _ in range(20):
            fn()
        t = []
        for _ in range(args.calls):
            t0 = time.perf_counter()
            fn()
            t.append(time.perf_counter() - t0)
        t = np.array(t) * 1e6
        return {"median_us": float(np.median(t)), "p10_us": float(np.percentile(t, 10)), "p90_us": float(np.percentile(t, 90))}

    us = np.array([5], np.int32)
    out["recommend_top10"] = timed(lambda: m.recommend(us, 10))
    for nq in (1, 3, 5, 8):
        q = rng.integers(0, ni, nq).astype(np.int32)
        out[f"similar_top20_nq{nq}"] = timed(lambda: m.similar(q, 20))
    out["hbm_floor_us"] = ni * k * 4 / 6.577e12 * 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
