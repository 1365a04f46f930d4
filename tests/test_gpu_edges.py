"""Edge cases of the LQR entry points on the device: a single node without edges (the
recursion degenerates to the root solve, lqr.cpp:798-819) and a batch of one on every kernel
family, against the CPU oracle."""
import numpy as np
import pytest

import problem_gen as pg
from gpu_helpers import REL_TOL, assert_lqr_parity, gpu_lqr_factor_solve
from oracle import pyoracle
from oracle.pyoracle import Structure

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [3, 12])
def test_single_node_without_edges(n):
    s = Structure.chain(0, n, 1)
    rng = np.random.default_rng(n)
    batch = 5
    Z = rng.standard_normal((batch, n, n))
    Q = np.einsum("bkj,bki->bij", Z, Z) + 1e-3 * np.eye(n)
    none = np.zeros((batch, 0))
    host = dict(Q=np.ascontiguousarray(np.swapaxes(Q, -1, -2)).reshape(batch, -1), M=none, R=none,
                q=rng.standard_normal((batch, n)), r=none, A=none, B=none,
                c=rng.standard_normal((batch, n)), delta=1e-3 + 0.1 * rng.random((batch, n)))
    ref = pyoracle.lqr_factor_solve(s, host)
    gpu, _ = gpu_lqr_factor_solve(s, host)
    assert (gpu["status"] == 0).all()
    assert_lqr_parity(gpu, ref, REL_TOL)


@pytest.mark.parametrize("n,m,T", [(12, 4, 50), (4, 1, 100), (64, 24, 4), (3, 2, 7), (12, 4, 1)])
def test_batch_of_one(n, m, T):
    s, host = pg.lqr_benchmark_batch(n, m, T, 1, seed=1)
    ref = pyoracle.lqr_factor_solve(s, host)
    gpu, _ = gpu_lqr_factor_solve(s, host)
    assert (gpu["status"] == 0).all()
    assert_lqr_parity(gpu, ref, REL_TOL)
