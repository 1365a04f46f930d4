"""Optional FP32 mode (sipoc_lqr_factor_solve_f32, csrc/riccati_f32.cu) against the FP64 CPU
oracle.

Stated tolerance of the mode: 5e-4 relative on (x, u, y) on the reference benchmark's
distribution (delta in [1e-3, 0.101]), the oracle being run on the SAME float-rounded inputs
(so the comparison measures the single-precision arithmetic, not the rounding of the data).
The same kernels instantiated on double are held to the FP64 tolerance (1e-9): that pins the
algebra, the FP32 number then only reflects the precision."""
import numpy as np
import pytest

import problem_gen as pg
from gpu_helpers import REL_TOL, assert_lqr_parity, rel_err, to_structs
from oracle import pyoracle
from sip_optimal_control_b200 import LQR

pytestmark = pytest.mark.gpu

FP32_TOL = 5e-4


def _round_to_f32(host):
    return {k: v.astype(np.float32).astype(np.float64) for k, v in host.items()}


@pytest.mark.parametrize("m", [1, 2, 3, 4])
@pytest.mark.parametrize("T", [1, 16, 100])
def test_fp32_mode_against_oracle(m, T):
    n, batch = 4, 77
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=10 * m + T, dense_M=True)
    host = _round_to_f32(host)
    ref = pyoracle.lqr_factor_solve(s, host)
    assert (ref["status"] == 0).all()
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, batch)
    assert lqr.f32_supported
    inp = lqr.pack_input(host)
    # single precision
    out32 = lqr.alloc_output_f32()
    st = lqr.factor_solve_f32(lqr.narrow_f32(inp), out32)
    assert (st[:batch].cpu().numpy() == 0).all()
    gpu32 = lqr.unpack_output({k: v.double() for k, v in out32.items()})
    worst = max(rel_err(gpu32[k], ref[k]).max() for k in ("x", "u", "y"))
    print(f"n=4 m={m} T={T}: FP32 vs FP64 oracle {worst:.2e}")
    assert worst < FP32_TOL
    # the same kernels on double: the FP64 tolerance
    out64 = lqr.alloc_output()
    st = lqr.factor_solve_thread_f64(inp, out64)
    assert (st[:batch].cpu().numpy() == 0).all()
    gpu64 = lqr.unpack_output(out64)
    gpu64["status"] = st[:batch].cpu().numpy()
    assert_lqr_parity(gpu64, ref, REL_TOL)


def test_fp32_mode_reports_failures_per_problem():
    n, m, T, batch = 4, 2, 20, 9
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=4)
    host = _round_to_f32(host)
    host["delta"][2, 7 * n + 1] = -0.5                                      # INVALID_DELTA
    host["R"][5, 11 * m * m:12 * m * m] = (-3.0 * np.eye(m)).flatten()      # G fails at edge 11
    ref = pyoracle.lqr_factor_solve(s, host)
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, batch)
    out32 = lqr.alloc_output_f32()
    st = lqr.factor_solve_f32(lqr.narrow_f32(lqr.pack_input(host)), out32)[:batch].cpu().numpy()
    assert (st == ref["status"]).all(), (st, ref["status"])
    good = ref["status"] == 0
    gpu32 = lqr.unpack_output({k: v.double() for k, v in out32.items()})
    for k in ("x", "u", "y"):
        assert rel_err(gpu32[k][good], ref[k][good]).max() < FP32_TOL


def test_fp32_mode_refuses_other_shapes():
    s, host = pg.lqr_benchmark_batch(12, 4, 5, 3, seed=0)
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, 3)
    assert not lqr.f32_supported
    with pytest.raises(Exception):
        lqr.factor_solve_f32(lqr.narrow_f32(lqr.pack_input(host)), lqr.alloc_output_f32())
