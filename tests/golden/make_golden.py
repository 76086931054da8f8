"""Regression fixtures for the ALS path.  There are no golden vectors for this path in the reference (the arithmetic
lives in Spark MLlib, which is not under /root/reference - SURVEY.md section 8(c)), so these are NOT reference outputs:
they are produced by the dense NumPy/SciPy restatement of the algorithm (oracle/als_oracle.py: numpy_als_train, built on
scipy.linalg.cho_factor / cho_solve in fp64 - independent of the C oracle's packed dspr/dppsv code and of the CUDA
kernels) on small seeded problems, and pinned here so that the C oracle, the CUDA path and future refactors are all
checked against the same stored numbers.

    python tests/golden/make_golden.py        # rewrites tests/golden/als_small.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import pio_b200  # noqa: E402,F401
from pio_b200 import synth  # noqa: E402
from oracle import als_oracle as o  # noqa: E402

CASES = [
    # name, users, items, ratings, rank, iterations, lambda, implicit, alpha
    ("explicit_r10", 300, 80, 4000, 10, 5, 0.01, False, 1.0),      # the recommendation template's defaults, scaled down
    ("implicit_r64", 400, 60, 6000, 64, 3, 0.05, True, 1.0),       # the headline configuration, scaled down
    ("implicit_r7_alpha", 200, 50, 3000, 7, 4, 0.1, True, 2.5),
]


def main():
    out = {}
    for name, nu, ni, nnz, rank, iters, lam, implicit, alpha in CASES:
        u, i, r = synth.synth_ratings(nu, ni, nnz, seed=17, implicit=implicit)
        u0 = synth.synth_init_factors(nu, rank, 9, 0)
        i0 = synth.synth_init_factors(ni, rank, 9, 1)
        uf, itf, uh, ih = o.numpy_als_train(nu, ni, u, i, r, rank, iters, lam, implicit, alpha, u0, i0)
        out[name + "/user"] = u
        out[name + "/item"] = i
        out[name + "/rating"] = r
        out[name + "/user_init"] = u0
        out[name + "/item_init"] = i0
        out[name + "/user_factors"] = np.asarray(uf, np.float64)
        out[name + "/item_factors"] = np.asarray(itf, np.float64)
        out[name + "/user_has"] = np.asarray(uh, np.uint8)
        out[name + "/item_has"] = np.asarray(ih, np.uint8)
        out[name + "/params"] = np.array([nu, ni, rank, iters, lam, float(implicit), alpha], np.float64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "als_small.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
