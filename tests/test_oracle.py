"""CPU tests of the oracle itself: the C restatement is checked against an
independent dense NumPy/SciPy restatement, a closed form, and the mathematical invariants,
because the reference holds no golden vectors for ALS (parity unpinned)."""
import numpy as np
import pytest

from pio_b200 import synth


def _problem(nu, ni, nnz, seed, implicit):
    u, i, r = synth.synth_ratings(nu, ni, nnz, seed=seed, implicit=implicit)
    return u, i, r


@pytest.mark.parametrize("implicit", [False, True])
@pytest.mark.parametrize("rank", [1, 3, 10])
def test_c_oracle_matches_dense_numpy(oracle, implicit, rank):
    nu, ni, nnz = 40, 30, 300
    u, i, r = _problem(nu, ni, nnz, 7, implicit)
    if implicit:
        r = r.copy()
        r[::7] = -r[::7]       # negative preferences: confidence only, no b term
        r[::11] = 0.0
    u0 = synth.synth_init_factors(nu, rank, 5, 0)
    i0 = synth.synth_init_factors(ni, rank, 5, 1)
    a = oracle.als_train(nu, ni, u, i, r, rank, 3, 0.05, implicit, 1.5, u0, i0)
    b = oracle.numpy_als_train(nu, ni, u, i, r, rank, 3, 0.05, implicit, 1.5, u0, i0)
    for x, y in zip(a[:2], b[:2]):
        assert np.abs(x - y).max() <= 2e-6 * max(1.0, np.abs(y).max())
    assert (a[2] == b[2]).all() and (a[3] == b[3]).all()


def test_rank1_closed_form(oracle):
    # one user, explicit, rank 1: x = sum(r*y) / (sum(y^2) + lambda*n)
    items = np.arange(5, dtype=np.int32)
    users = np.zeros(5, np.int32)
    r = np.array([1, 2, 3, 4, 5], np.float32)
    i0 = np.array([[0.5], [-1.0], [2.0], [0.25], [1.5]], np.float32)
    ptr, idx, val = oracle.csr_build(1, users, items, r)
    dst = np.zeros((1, 1), np.float32)
    oracle.half_step(ptr, idx, val, i0, dst, 0.1, False, 1.0)
    y = i0[:, 0].astype(np.float64)
    expect = (r * y).sum() / ((y * y).sum() + 0.1 * 5)
    assert abs(dst[0, 0] - expect) < 1e-6


@pytest.mark.parametrize("implicit", [False, True])
def test_objective_is_monotone(oracle, implicit):
    nu, ni, nnz, k = 60, 25, 500, 4
    u, i, r = _problem(nu, ni, nnz, 11, implicit)
    uf = synth.synth_init_factors(nu, k, 2, 0)
    itf = synth.synth_init_factors(ni, k, 2, 1)
    prev = None
    for it in range(5):
        uf, itf, uh, ih = oracle.als_train(nu, ni, u, i, r, k, 1, 0.1, implicit, 1.0, uf, itf)
        obj = oracle.als_objective(u, i, r, uf, itf, 0.1, implicit, 1.0)
        if prev is not None:
            assert obj <= prev * (1 + 1e-6)
        prev = obj


def test_rows_without_ratings_get_no_factor(oracle):
    u = np.array([0, 0, 2], np.int32)
    i = np.array([1, 3, 1], np.int32)
    r = np.array([5, 3, 1], np.float32)
    u0 = synth.synth_init_factors(4, 2, 1, 0)
    i0 = synth.synth_init_factors(5, 2, 1, 1)
    uf, itf, uh, ih = oracle.als_train(4, 5, u, i, r, 2, 2, 0.01, False, 1.0, u0, i0)
    assert uh.tolist() == [1, 0, 1, 0] and ih.tolist() == [0, 1, 0, 1, 0]
    assert (uf[1] == 0).all() and (uf[3] == 0).all() and (itf[0] == 0).all()


def test_duplicates_are_separate_ratings(oracle):
    # recommendation template does not dedup (ALSAlgorithm.scala:62-65): a repeated pair counts twice
    i0 = np.array([[1.0], [2.0]], np.float32)
    dst = np.zeros((1, 1), np.float32)
    ptr, idx, val = oracle.csr_build(1, np.zeros(3, np.int32), np.array([0, 0, 1], np.int32),
                                     np.array([4, 2, 1], np.float32))
    oracle.half_step(ptr, idx, val, i0, dst, 0.5, False, 1.0)
    expect = (4 * 1 + 2 * 1 + 1 * 2) / (1 + 1 + 4 + 0.5 * 3)
    assert abs(dst[0, 0] - expect) < 1e-6


def test_dedup_modes(oracle):
    u = np.array([1, 0, 1, 1, 0], np.int32)
    i = np.array([2, 0, 2, 2, 0], np.int32)
    r = np.array([1, 2, 3, 4, 5], np.float32)
    ts = np.array([10, 5, 30, 20, 5], np.int64)
    du, di, dr = oracle.dedup_coo(u, i, r, "sum")
    assert du.tolist() == [0, 1] and di.tolist() == [0, 2] and dr.tolist() == [7.0, 8.0]
    du, di, dr = oracle.dedup_coo(u, i, r, "keep_last", ts)
    assert dr.tolist() == [5.0, 3.0]  # (0,0): equal ts -> later event; (1,2): ts=30
    du, di, dr = oracle.dedup_coo(u, i, r, "none")
    assert du.tolist() == [0, 0, 1, 1, 1] and dr.tolist() == [2, 5, 1, 3, 4]


def test_topk_and_similar(oracle):
    rng = np.random.default_rng(0)
    uf = rng.standard_normal((6, 5)).astype(np.float32)
    itf = rng.standard_normal((40, 5)).astype(np.float32)
    ih = np.ones(40, np.uint8)
    ih[3] = 0
    mask = np.zeros(40, np.uint8)
    mask[7] = 1
    items, scores, cnt = oracle.recommend(uf, None, itf, ih, np.array([0, 5], np.int32), 4, mask)
    full = uf[[0, 5]].astype(np.float64) @ itf.astype(np.float64).T
    full[:, 3] = -np.inf
    full[:, 7] = -np.inf
    for q in range(2):
        order = np.argsort(-full[q], kind="stable")[:4]
        assert items[q].tolist() == order.tolist()
        assert np.allclose(scores[q], full[q][order], rtol=1e-6)
    q = np.array([1, 2], np.int32)
    si, ss, sc = oracle.similar(itf, ih, q, 5, mask)
    f64 = itf.astype(np.float64)
    nrm = np.linalg.norm(f64, axis=1)
    cos = (f64 @ f64[q].T) / (nrm[:, None] * nrm[q][None, :])
    s = cos.sum(1)
    s[[1, 2, 3, 7]] = -np.inf
    s[s <= 0] = -np.inf
    order = np.argsort(-s, kind="stable")[:5]
    assert si.tolist() == order.tolist()


def test_naive_bayes_formulas(oracle):
    x = np.array([[1, 0, 2], [0, 3, 1], [2, 2, 0], [1, 1, 1]], np.float32)
    y = np.array([0, 1, 0, 1], np.int32)
    pi, theta = oracle.nb_train(y, x, 2, 1.0)
    assert np.allclose(pi, np.log([3 / 6, 3 / 6]))
    s0 = np.array([3, 2, 2], float)
    assert np.allclose(theta[0], np.log((s0 + 1) / (s0.sum() + 3)))
    scores = pi[None, :] + x.astype(np.float64) @ theta.T
    assert oracle.nb_predict(x, pi, theta).tolist() == scores.argmax(1).tolist()


def test_low_rank_recovery_cases(oracle):
    """MLlib-suite style recovery checks (tests/lowrank_cases.py) on the fp64-accumulating oracle."""
    import lowrank_cases as L
    for name, nu, ni, rank, noise, iters, reg, implicit, metric, target in L.CASES:
        tr, te = L.gen(nu, ni, rank, noise, implicit)
        u0 = synth.synth_init_factors(nu, rank, 5, 0)
        i0 = synth.synth_init_factors(ni, rank, 5, 1)
        uf, itf, hu, hi = oracle.als_train(nu, ni, tr[0], tr[1], tr[2], rank, iters, reg, implicit, 1.0, u0, i0)
        v = L.score(metric, uf, itf, te, hu, hi)
        assert L.passes(metric, v, target), (name, metric, v, target)


def _golden_cases():
    from pathlib import Path
    z = np.load(Path(__file__).parent / "golden" / "als_small.npz")
    names = sorted({k.split("/")[0] for k in z.files})
    for n in names:
        nu, ni, rank, iters, lam, implicit, alpha = z[n + "/params"]
        yield n, z, int(nu), int(ni), int(rank), int(iters), float(lam), bool(implicit), float(alpha)


def test_c_oracle_matches_stored_fixtures(oracle):
    """tests/golden/als_small.npz (dense NumPy/SciPy restatement, see make_golden.py): the C oracle reproduces the stored
    factors to fp32 rounding of the stored factors between half-steps."""
    for n, z, nu, ni, rank, iters, lam, implicit, alpha in _golden_cases():
        uf, itf, hu, hi = oracle.als_train(nu, ni, z[n + "/user"], z[n + "/item"], z[n + "/rating"], rank, iters, lam,
                                           implicit, alpha, z[n + "/user_init"], z[n + "/item_init"])
        assert np.array_equal(hu, z[n + "/user_has"]) and np.array_equal(hi, z[n + "/item_has"])
        for got, want in ((uf, z[n + "/user_factors"]), (itf, z[n + "/item_factors"])):
            err = np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want)
            assert err <= 2e-6, (n, err)


def test_csr_dedup_sum_equals_dedup_coo(oracle):
    """oracle_csr_dedup_sum (used by bench.py's parity sample at 100 M ratings) == dedup_coo(mode="sum")."""
    rng = np.random.default_rng(0)
    nu, ni, nnz = 300, 40, 6000
    u = rng.integers(0, nu, nnz).astype(np.int32)
    i = rng.integers(0, ni, nnz).astype(np.int32)
    r = (rng.integers(1, 9, nnz) + rng.random(nnz)).astype(np.float32)
    uu, ii, rr = oracle.dedup_coo(u, i, r, "sum")
    ptr, col, val = oracle.csr_dedup_sum(*oracle.csr_build(nu, u, i, r))
    assert np.array_equal(oracle.csr_rows(ptr), uu) and np.array_equal(col, ii) and np.array_equal(val, rr)
    assert ptr[-1] == rr.shape[0] < nnz


@pytest.mark.parametrize("implicit", [False, True])
def test_half_step_rows_equals_half_step(oracle, implicit):
    rng = np.random.default_rng(1)
    nu, ni, nnz, k = 200, 50, 3000, 8
    u, i, r = _problem(nu, ni, nnz, 9, implicit)
    ptr, col, val = oracle.csr_build(nu, u, i, r)
    src = rng.standard_normal((ni, k)).astype(np.float32)
    dst = np.zeros((nu, k), np.float32)
    yty = oracle.gram(src) if implicit else None
    assert oracle.half_step(ptr, col, val, src, dst, 0.05, implicit, 1.0, yty) == 0
    rows = np.array([0, 5, 17, 199], np.int32)
    out, fails = oracle.half_step_rows(ptr, col, val, src, rows, 0.05, implicit, 1.0, yty)
    assert fails == 0 and np.array_equal(out, dst[rows])
