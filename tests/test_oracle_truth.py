"""The FP64 oracle against an extended-precision (80-bit) statement of the same recursion
(tests/extended_precision.py): how far the Eigen-free port is from the exact answer on the
benchmark distribution and at the Newton-KKT regularization range (delta up to 1e9), where
agreement between two FP64 implementations alone would not show it."""
import numpy as np
import pytest

import extended_precision as xp
import problem_gen as pg
from oracle import pyoracle


@pytest.mark.parametrize("n,m,T", [(12, 4, 12), (6, 2, 20), (4, 1, 30)])
@pytest.mark.parametrize("wide", [False, True])
def test_port_error_against_extended_precision(n, m, T, wide):
    s, host = pg.lqr_benchmark_batch(n, m, T, 3, seed=n + T, dense_M=True)
    if wide:
        host = xp.wide_delta(host, seed=T)
    truth = xp.lqr_chain_truth_batch(n, m, T, host)
    ref = pyoracle.lqr_factor_solve(s, host)
    assert (ref["status"] == 0).all()
    err = xp.error_against(truth, ref)
    print(f"n={n} m={m} T={T} wide={wide}: port vs 80-bit truth {err:.2e}")
    assert err < 1e-9  # north_star's tolerance, against the truth rather than a sibling


def test_truth_solves_the_kkt_system():
    # the extended-precision recursion satisfies the reference's residual (lqr_test.cpp:157-180)
    n, m, T = 6, 2, 10
    s, host = pg.lqr_benchmark_batch(n, m, T, 2, seed=5, dense_M=True)
    truth = xp.lqr_chain_truth_batch(n, m, T, host)
    res = pyoracle.lqr_residual(s, host, truth["x"], truth["u"], truth["y"])
    scale = np.linalg.norm(np.concatenate([host["q"], host["r"], host["c"]], axis=1), axis=1)
    assert (res / scale).max() < 1e-13
