"""GPU test (-m gpu) of the lockstep warp Cholesky (csrc/als_lockstep.cuh) in isolation: dense SPD systems through the
debug entry pio_als_debug_lockstep against numpy's fp64 solve.  The routine replaces MLlib's CholeskySolver.solve
 inside every half-step kernel of rank 33..64 (two matrices per warp) and rank 65..128 (one)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_lockstep(native, A, b, ridge, reps=1):
    n, N, _ = A.shape
    A = np.ascontiguousarray(A, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    x = np.zeros((n, N), np.float32)
    ms = C.c_float(0)
    fail = C.c_int(0)
    f = native.lib().pio_als_debug_lockstep
    rc = f(C.c_int(0), C.c_int(N), C.c_int(n), A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
           C.c_float(ridge), x.ctypes.data_as(C.c_void_p), C.c_int(reps), C.byref(ms), C.byref(fail))
    assert rc == 0, native.lib().pio_als_last_error(None)
    return x, ms.value, fail.value


@pytest.mark.parametrize("N", [64, 128])
@pytest.mark.parametrize("n", [1, 2, 3, 257])
def test_lockstep_solver_matches_numpy(native, N, n):
    rng = np.random.default_rng(N + n)
    Y = rng.standard_normal((n, 3 * N, N)).astype(np.float32)
    A = np.einsum("nri,nrj->nij", Y, Y).astype(np.float32)
    b = rng.standard_normal((n, N)).astype(np.float32)
    x, _, fail = run_lockstep(native, A, b, 0.25)
    assert fail == 0
    for m in range(n):
        ref = np.linalg.solve(A[m].astype(np.float64) + 0.25 * np.eye(N), b[m].astype(np.float64))
        assert np.abs(x[m] - ref).max() <= 2e-5 * np.abs(ref).max(), (m, np.abs(x[m] - ref).max(), np.abs(ref).max())


def test_lockstep_solver_flags_indefinite_matrices(native):
    N = 64
    A = np.zeros((2, N, N), np.float32)
    A[0] = np.eye(N)
    A[1] = -np.eye(N)
    b = np.ones((2, N), np.float32)
    x, _, fail = run_lockstep(native, A, b, 0.0)
    assert fail == 1 and np.allclose(x[0], 1.0)
