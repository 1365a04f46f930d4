"""GPU kernels against the extended-precision truth (tests/extended_precision.py), next to the
CPU oracle's own error against it: at delta up to 1e9 (the Newton-KKT benchmark's r2 range)
the default shape-specialised plans are as close to the exact answer as the reference-order
port is -- which bounds their distance to the reference's Eigen path too (any backward-stable
FP64 statement of the recursion sits within the same ~1e-13 of the truth)."""
import pytest

import extended_precision as xp
import problem_gen as pg
from gpu_helpers import gpu_lqr_factor_solve
from oracle import pyoracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,T", [(12, 4, 12), (6, 2, 20), (4, 1, 30), (8, 3, 10)])
@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("fused", [True, False])
def test_default_plans_against_extended_precision(n, m, T, wide, fused):
    s, host = pg.lqr_benchmark_batch(n, m, T, 3, seed=n + T, dense_M=True)
    if wide:
        host = xp.wide_delta(host, seed=T)
    truth = xp.lqr_chain_truth_batch(n, m, T, host)
    ref = pyoracle.lqr_factor_solve(s, host)
    gpu, lqr = gpu_lqr_factor_solve(s, host, fused=fused)
    assert (gpu["status"] == 0).all()
    assert "generic" not in lqr.engine.kernel_variant
    e_port, e_gpu = xp.error_against(truth, ref), xp.error_against(truth, gpu)
    print(f"{lqr.engine.kernel_variant} wide={wide} fused={fused}: port {e_port:.2e} gpu {e_gpu:.2e}")
    assert e_gpu < 1e-9                       # the tolerance, against the truth
    assert e_gpu < 50 * e_port + 1e-13        # and of the order of the port's own error


def test_scan_against_extended_precision():
    n, m, T = 12, 4, 64
    s, host = pg.lqr_benchmark_batch(n, m, T, 2, seed=3, dense_M=True)
    truth = xp.lqr_chain_truth_batch(n, m, T, host)
    gpu, lqr = gpu_lqr_factor_solve(s, host, parallel_in_time=True)
    assert lqr.engine.kernel_variant.startswith("scan_")
    assert xp.error_against(truth, gpu) < 1e-9
