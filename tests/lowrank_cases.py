"""Low-rank recovery cases in the style of Spark MLlib's own ALS test-suite (ml/recommendation/ALSSuite.scala:
"exact rank-1 matrix", "approximate rank-1 / rank-2 matrix", "implicit feedback"): ratings generated from random
low-rank factors (+ Gaussian noise), a random 60 % of the entries used for training, and the error of the predictions
on the held-out entries compared with a target.  MLlib is a third-party dependency that is not under /root/reference,
so nothing here is a golden vector: generator, iteration counts and targets are this repository's own (chosen with
margin from the fp64 oracle's behaviour); the cases check that both the oracle and the CUDA path actually recover the
structure MLlib's suite checks for, on top of the per-row parity tests."""
import numpy as np

CASES = [
    # name,                     nu, ni, rank, noise, iters, reg,  implicit, metric, target
    ("exact rank-1 matrix", 20, 40, 1, 0.0, 8, 1e-5, False, "rmse", 1e-3),
    ("approximate rank-1 matrix", 20, 40, 1, 0.01, 12, 0.01, False, "rmse", 0.03),
    ("approximate rank-2 matrix", 20, 40, 2, 0.01, 12, 0.01, False, "rmse", 0.05),
    ("implicit feedback", 20, 40, 2, 0.01, 8, 0.01, True, "auc", 0.75),
]


def gen(nu, ni, rank, noise, implicit, seed=11, train_frac=0.6):
    """Factors uniform in [0, 1)/sqrt(rank)-scaled (so ratings are O(1)), like ALSSuite.genFactors; for implicit
    feedback the observed value is the confidence-style count and the target preference is 1 where it is positive."""
    rng = np.random.default_rng(seed)
    a = (1.0 / np.sqrt(rank))
    uf = rng.uniform(-a, a, size=(nu, rank))
    vf = rng.uniform(-a, a, size=(ni, rank))
    full = uf @ vf.T
    uu, ii = np.meshgrid(np.arange(nu), np.arange(ni), indexing="ij")
    uu, ii, rr = uu.ravel(), ii.ravel(), full.ravel()
    if implicit:
        # ALSSuite.genImplicitTestData: observed rating = scaled positive part, truth = 1 if rating > 0 else 0
        truth = (rr > 0).astype(np.float64)
        obs = np.where(rr > 0, 1.0 + 4.0 * rr / max(rr.max(), 1e-9), 0.0)
    else:
        truth = rr
        obs = rr
    obs = obs + noise * rng.standard_normal(obs.shape)
    mask = rng.uniform(size=obs.shape) < train_frac
    tr = (uu[mask].astype(np.int32), ii[mask].astype(np.int32), obs[mask].astype(np.float32))
    te = (uu[~mask], ii[~mask], truth[~mask])
    return tr, te


def rmse(uf, itf, te, has_u=None, has_i=None):
    u, i, t = te
    ok = np.ones(len(u), bool)
    if has_u is not None:
        ok &= has_u[u].astype(bool) & has_i[i].astype(bool)
    pred = np.einsum("ij,ij->i", uf[u[ok]].astype(np.float64), itf[i[ok]].astype(np.float64))
    return float(np.sqrt(np.mean((pred - t[ok]) ** 2)))


def auc(uf, itf, te):
    """Probability that a held-out positive preference is scored above a held-out zero preference."""
    u, i, t = te
    pred = np.einsum("ij,ij->i", uf[u].astype(np.float64), itf[i].astype(np.float64))
    pos, neg = pred[t > 0], pred[t == 0]
    return float((pos[:, None] > neg[None, :]).mean())


def score(metric, uf, itf, te, has_u, has_i):
    return rmse(uf, itf, te, has_u, has_i) if metric == "rmse" else auc(uf, itf, te)


def passes(metric, value, target):
    return value <= target if metric == "rmse" else value >= target
