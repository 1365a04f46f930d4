"""Parallel-in-time factor + solve (csrc/scan.cu + the segmented sweep of riccati_fast.cu)
against the serial CPU oracle.  Tolerance: the scan combines conditional value functions of
whole segments, so it agrees with the serial recursion (lqr.cpp:645-871) to rounding times
the conditioning of those combinations rather than operation by operation -- 1e-9 relative
on (x, u, y) at the benchmark regularization range still holds and is what is asserted."""
import numpy as np
import pytest

import problem_gen as pg
from gpu_helpers import REL_TOL, assert_lqr_parity, gpu_lqr_factor_solve
from oracle import pyoracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,T", [(12, 4, 64), (6, 1, 96), (6, 2, 128), (8, 3, 60), (12, 4, 6),
                                   (6, 4, 40), (8, 1, 48)])
def test_forced_scan_matches_oracle(n, m, T):
    batch = 19
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=T + n, dense_M=True)
    ref = pyoracle.lqr_factor_solve(s, host)
    assert (ref["status"] == 0).all()
    gpu, lqr = gpu_lqr_factor_solve(s, host, parallel_in_time=True)
    assert lqr.engine.kernel_variant.startswith("scan_"), lqr.engine.kernel_variant
    assert (gpu["status"] == 0).all()
    assert_lqr_parity(gpu, ref, REL_TOL)
    scale = np.linalg.norm(np.concatenate([host["q"], host["r"], host["c"]], axis=1), axis=1)
    assert (gpu["residual"] / scale).max() < 1e-9
    # ... and against the serial kernels of the same engine
    ser, lqr2 = gpu_lqr_factor_solve(s, host, parallel_in_time=False)
    assert not lqr2.engine.kernel_variant.startswith("scan_")
    assert_lqr_parity(gpu, ser, 1e-10)


@pytest.mark.parametrize("n,m,T,batch", [(12, 4, 4096, 4), (6, 2, 2048, 9)])
def test_long_horizon_picks_the_scan_by_itself(n, m, T, batch):
    # BASELINE config 5b: N = 4 096 at quadrotor dims (batch 64 in the bench)
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=7)
    ref = pyoracle.lqr_factor_solve(s, host)
    assert (ref["status"] == 0).all()
    gpu, lqr = gpu_lqr_factor_solve(s, host)
    assert lqr.engine.kernel_variant.startswith("scan_"), lqr.engine.kernel_variant
    assert (gpu["status"] == 0).all()
    assert_lqr_parity(gpu, ref, REL_TOL)


def test_scan_reports_failures_per_problem():
    n, m, T, batch = 6, 2, 48, 11
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=2)
    host["delta"][3, 17 * n + 1] = -0.5           # INVALID_DELTA at node 17
    host["Q"][7, 30 * n * n:31 * n * n] = (-40.0 * np.eye(n)).flatten()  # F fails at node 30
    ref = pyoracle.lqr_factor_solve(s, host)
    gpu, lqr = gpu_lqr_factor_solve(s, host, parallel_in_time=True)
    assert lqr.engine.kernel_variant.startswith("scan_")
    good = ref["status"] == 0
    assert good.sum() == batch - 2
    assert ((gpu["status"] == 0) == good).all()
    assert gpu["status"][3] == 1                   # the delta check is the sweep's own
    assert_lqr_parity(gpu, ref, REL_TOL, mask=good)
