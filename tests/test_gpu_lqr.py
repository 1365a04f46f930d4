"""GPU parity tests of the LQR path, all through the C ABI (libsipoc.so).

Every case restates a reference test (tests/lqr_test.cpp) or a reference
benchmark configuration and compares the CUDA result with the CPU oracle on the
same inputs.  Tolerance: 1e-9 relative per problem on x, u, y (north_star); the
reference's own residual bars (1e-12) are checked as well.
"""
import numpy as np
import pytest

import problem_gen as pg
import reference_fixtures as fx
from gpu_helpers import REL_TOL, assert_lqr_parity, gpu_lqr_factor_solve, rel_err, to_structs
from oracle import pyoracle
from oracle.pyoracle import Structure
from sip_optimal_control_b200 import LQR, Dimensions, Topology, FactorStatus

pytestmark = pytest.mark.gpu


def _tile(one: dict, batch: int) -> dict:
    return {k: np.repeat(v, batch, axis=0) for k, v in one.items()}


# --- reference fixtures ---------------------------------------------------------
@pytest.mark.parametrize("builder", [fx.nonuniform_delta_chain, fx.branch_tree,
                                     fx.variable_dim_branch_tree, fx.five_node_tree])
@pytest.mark.parametrize("fused", [True, False])
def test_reference_fixtures_match_oracle(builder, fused):
    # lqr_test.cpp:229-263, :411-429, :641-659, :982-1013
    s, p = builder()
    host = _tile(fx.pack_problem(**p), 3)
    ref = pyoracle.lqr_factor_solve(s, host)
    gpu, lqr = gpu_lqr_factor_solve(s, host, fused=fused)
    assert (gpu["status"] == 0).all()
    assert_lqr_parity(gpu, ref, 1e-12)
    assert gpu["residual"].max() < 1e-12          # the reference's own bar
    # the oracle's residual function agrees with the device one
    r = pyoracle.lqr_residual(s, host, gpu["x"], gpu["u"], gpu["y"])
    assert np.abs(r - gpu["residual"]).max() < 1e-13


def test_matches_dense_kkt_on_five_node_tree():  # lqr_test.cpp:982-1013
    s, p = fx.five_node_tree()
    gpu, _ = gpu_lqr_factor_solve(s, fx.pack_problem(**p))
    xs, us, ys = fx.dense_kkt_solve(s, p)
    assert rel_err(gpu["x"], np.concatenate(xs)[None])[0] < 1e-10
    assert rel_err(gpu["u"], np.concatenate(us)[None])[0] < 1e-10
    assert rel_err(gpu["y"], np.concatenate(ys)[None])[0] < 1e-10


def test_compiled_topology_orders():  # lqr_test.cpp:931-953
    s, _ = fx.five_node_tree()
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, 2)
    co, ce, pre, post = lqr.engine.compiled_topology()
    assert co.tolist() == [0, 2, 4, 4, 4, 4]
    assert ce.tolist() == [0, 1, 2, 3]
    assert pre.tolist() == [0, 1, 3, 4, 2]
    assert post.tolist() == [2, 4, 3, 1, 0]


def test_invalid_topology_is_latched():  # lqr_test.cpp:452-464
    s = Structure([0, 0], [1, 1], 0, [2, 2, 2], [1, 1])
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, 2)
    assert lqr.traversal_status_ == FactorStatus.INVALID_TOPOLOGY
    st = lqr.factor_with_status({})
    assert (np.asarray(st) == FactorStatus.INVALID_TOPOLOGY).all()
    assert not np.asarray(lqr.factor({})).any()


# --- status codes (lqr_test.cpp:188-227), mixed inside one batch ----------------
@pytest.mark.parametrize("force_generic", [True, False])
def test_status_codes_per_problem(force_generic):
    n, m, T = 2, 1, 2
    s, p = fx.identity_chain(n, m, T)
    host = _tile(fx.pack_problem(**p), 6)
    sz = pyoracle.lqr_sizes(s)
    # problem 1: delta[T][0] = 0 -> INVALID_DELTA (lqr_test.cpp:206-211)
    host["delta"][1, T * n + 0] = 0.0
    # problem 2: Q[T] = -2 I -> F failure (cf. lqr_test.cpp:213-219)
    host["Q"][2, T * n * n:] = (-2.0 * np.eye(n)).flatten()
    # problem 3: Q[T] = 0, R[T-1] = -1 -> G failure (cf. lqr_test.cpp:221-227)
    host["Q"][3, T * n * n:] = 0.0
    host["R"][3, (T - 1) * m * m] = -1.0
    # problem 4: negative delta at the root AND a G failure deeper: post-order
    # visits the leaf side first, so G wins (lqr.cpp:696-700 before :722-727).
    host["delta"][4, 1] = -1.0
    host["R"][4, (T - 1) * m * m] = -50.0
    # problem 5 stays valid.
    ref = pyoracle.lqr_factor_solve(s, host)
    assert ref["status"].tolist() == [0, 1, 2, 3, 3, 0]
    gpu, _ = gpu_lqr_factor_solve(s, host, force_generic=force_generic)
    assert gpu["status"].tolist() == ref["status"].tolist()
    good = ref["status"] == 0
    assert_lqr_parity(gpu, ref, REL_TOL, mask=good)
    assert gpu["stats"][2] == 4 and gpu["stats"][3] == 6   # failed, count
    del sz


def test_single_stage_failure_fixtures():  # lqr_test.cpp:213-227 verbatim (n=m=T=1)
    s, p = fx.identity_chain(1, 1, 1)
    p["Q"][1][0, 0] = -2.0
    gpu, _ = gpu_lqr_factor_solve(s, fx.pack_problem(**p))
    assert gpu["status"][0] == FactorStatus.F_FACTORIZATION_FAILURE
    s, p = fx.identity_chain(1, 1, 1)
    p["Q"][1][0, 0] = 0.0
    p["R"][0][0, 0] = -1.0
    gpu, _ = gpu_lqr_factor_solve(s, fx.pack_problem(**p))
    assert gpu["status"][0] == FactorStatus.G_FACTORIZATION_FAILURE



@pytest.mark.parametrize("n,m,T", [(4, 1, 9), (12, 4, 6), (6, 2, 5), (16, 4, 4), (64, 24, 2),
                                   (3, 2, 5), (4, 2, 6), (2, 2, 4), (4, 3, 5)])
@pytest.mark.parametrize("fused", [True, False])
def test_status_codes_on_specialised_kernels(n, m, T, fused):
    """Failure injection on the shape-specialised paths (thread / sub-warp / CTA):
    same FactorStatus per problem as the oracle, healthy neighbours untouched."""
    batch = 11
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=5 + n)
    host = {k: v.copy() for k, v in host.items()}
    eye_n, eye_m = np.eye(n).flatten(), np.eye(m).flatten()
    host["delta"][1, T * n + 0] = 0.0                                  # INVALID_DELTA at node T
    host["Q"][2, T * n * n:] = -400.0 * eye_n                          # F failure at node T
    host["R"][3, (T - 1) * m * m:] = -50.0 * eye_m                     # G failure at edge T-1
    host["delta"][4, 1] = -1.0                                         # root delta AND ...
    host["R"][4, (T - 1) * m * m:] = -50.0 * eye_m                     # ... G deeper: G wins
    host["delta"][5, (T // 2) * n + n - 1] = -3.0                      # INVALID_DELTA mid-horizon
    host["Q"][6, 0:n * n] = -400.0 * eye_n                             # F failure at the root
    host["R"][7, 0:m * m] = -50.0 * eye_m                              # G failure at edge 0
    ref = pyoracle.lqr_factor_solve(s, host)
    assert ref["status"].tolist() == [0, 1, 2, 3, 3, 1, 2, 3, 0, 0, 0]
    gpu, lqr = gpu_lqr_factor_solve(s, host, fused=fused)
    assert "generic" not in lqr.engine.kernel_variant
    assert gpu["status"].tolist() == ref["status"].tolist()
    good = ref["status"] == 0
    assert_lqr_parity(gpu, ref, REL_TOL, mask=good)
    assert gpu["stats"][2] == 7 and gpu["stats"][3] == batch


def test_status_stats_and_kernel_profile():
    """sipoc_status_stats (the failure flags a Newton iteration all-reduces) and the
    per-kernel event profiler of the C ABI."""
    import ctypes

    import torch

    from sip_optimal_control_b200._capi import lib

    n, m, T, batch = 12, 4, 5, 70
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=3)
    host["delta"][9, 0] = -1.0
    host["delta"][40, n] = 0.0
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, batch)
    eng = lqr.engine
    inp, out = lqr.pack_input(host), lqr.alloc_output()
    lib.sipoc_profile_enable(eng._handle, 1)
    status = lqr.factor_solve(inp, out)
    stats = torch.zeros(4, dtype=torch.float64, device=status.device)
    eng._check(lib.sipoc_status_stats(eng._handle, status.data_ptr(), stats.data_ptr(),
                                      eng.stream_ptr()))
    assert stats.cpu().tolist() == [0.0, 0.0, 2.0, float(batch)]
    names = {}
    for i in range(lib.sipoc_profile_collect(eng._handle)):
        nm, ms, cnt = ctypes.c_char_p(), ctypes.c_double(), ctypes.c_int64()
        eng._check(lib.sipoc_profile_get(eng._handle, i, ctypes.byref(nm), ctypes.byref(ms),
                                         ctypes.byref(cnt)))
        names[nm.value.decode()] = (ms.value, cnt.value)
    # (factor + solve in one call: the sweep and the rollout are one kernel)
    assert set(names) == {"riccati_fused_subwarp", "status_stats_kernel"}
    assert all(ms > 0.0 and cnt == 1 for ms, cnt in names.values())
    lib.sipoc_profile_enable(eng._handle, 0)
    assert lib.sipoc_profile_collect(eng._handle) == 0

# --- benchmark-distribution chains (lqr_benchmark.cpp:61-96, :537-545) -----------
SHAPES = [(4, 1, 16), (4, 1, 100), (6, 2, 32), (8, 3, 16), (12, 4, 50), (16, 4, 16),
          (5, 2, 7), (3, 3, 4), (32, 8, 6), (64, 24, 4)]


@pytest.mark.parametrize("n,m,T", SHAPES)
@pytest.mark.parametrize("force_generic", [True, False])
def test_benchmark_chains_match_oracle(n, m, T, force_generic):
    batch = 37  # ragged: not a multiple of the warp / tile size
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=100 + n + m + T,
                                     dense_M=(T % 2 == 0))
    ref = pyoracle.lqr_factor_solve(s, host)
    assert (ref["status"] == 0).all()
    gpu, lqr = gpu_lqr_factor_solve(s, host, force_generic=force_generic)
    assert (gpu["status"] == 0).all()
    assert_lqr_parity(gpu, ref, REL_TOL)
    # residual parity: both are rounding noise relative to the data; compare
    # on the scale of the right-hand side.
    scale = np.linalg.norm(np.concatenate([host["q"], host["r"], host["c"]], axis=1), axis=1)
    assert (gpu["residual"] / scale).max() < 1e-9
    assert np.abs(gpu["residual"] - ref["residual"]).max() / scale.max() < 1e-9
    assert np.isclose(gpu["stats"][1], gpu["residual"].max())
    assert np.isclose(gpu["stats"][0], (gpu["residual"] ** 2).sum())
    assert gpu["stats"][3] == batch
    if not force_generic and (n, m) in ((4, 1), (12, 4), (6, 2), (8, 3), (16, 4), (32, 8), (64, 24)):
        assert "generic" not in lqr.engine.kernel_variant, lqr.engine.kernel_variant


@pytest.mark.parametrize("n,m,T,batch", [(64, 24, 32, 4), (12, 4, 50, 24), (4, 1, 100, 40)])
@pytest.mark.parametrize("fused", [True, False])
def test_baseline_horizons_match_oracle(n, m, T, batch, fused):
    # The BASELINE configs at their full horizon (humanoid T = 32, quadrotor T = 50,
    # cartpole T = 100) on a small batch, element-wise against the oracle.
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=n + T)
    ref = pyoracle.lqr_factor_solve(s, host)
    assert (ref["status"] == 0).all()
    gpu, lqr = gpu_lqr_factor_solve(s, host, fused=fused)
    assert "generic" not in lqr.engine.kernel_variant
    assert (gpu["status"] == 0).all()
    assert_lqr_parity(gpu, ref, REL_TOL)


@pytest.mark.parametrize("n", [4, 6, 8, 16])
@pytest.mark.parametrize("m", [1, 2, 3, 4])
def test_reference_benchmark_grid_runs_on_shape_specialised_kernels(n, m):
    # lqr_benchmark.cpp:537-545: T in {16, 32, 64, 128} x n in {4, 6, 8, 16} x m in {1, 2, 3, 4}.
    # Every uniform shape of the grid has a compiled plan (n = 16 with m < 4 through decoupled
    # padding to the (16, 4) CTA plan); fused factor + solve and factor-then-solve both checked.
    T, batch = 16, 29
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=11 * n + m, dense_M=True)
    ref = pyoracle.lqr_factor_solve(s, host)
    assert (ref["status"] == 0).all()
    for fused in (True, False):
        gpu, lqr = gpu_lqr_factor_solve(s, host, fused=fused)
        assert "generic" not in lqr.engine.kernel_variant, lqr.engine.kernel_variant
        assert (gpu["status"] == 0).all()
        assert_lqr_parity(gpu, ref, REL_TOL)


@pytest.mark.parametrize("n,m,T", [(10, 3, 9), (7, 4, 12), (24, 6, 5), (40, 20, 3)])
def test_other_uniform_shapes_are_padded_to_the_next_plan(n, m, T):
    batch = 13
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=n + m)
    host["delta"][3, n] = -1.0  # status codes survive the padding
    ref = pyoracle.lqr_factor_solve(s, host)
    gpu, lqr = gpu_lqr_factor_solve(s, host)
    assert lqr.engine.kernel_variant.startswith("padded_to_"), lqr.engine.kernel_variant
    assert (gpu["status"] == ref["status"]).all() and gpu["status"][3] == 1
    assert_lqr_parity(gpu, ref, REL_TOL, mask=ref["status"] == 0)


@pytest.mark.parametrize("n,m,T", [(4, 1, 20), (12, 4, 10), (5, 2, 7), (16, 4, 6), (64, 24, 3)])
def test_factor_once_solve_many(n, m, T):
    # BM_LQRSolve semantics (lqr_benchmark.cpp:611-638): re-solve with new
    # q, r, c against a kept factorization.
    batch = 19
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=7)
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, batch)
    inp = lqr.pack_input(host)
    out = lqr.alloc_output()
    status = lqr.factor_with_status(inp)
    assert (status[:batch].cpu().numpy() == 0).all()
    rng = np.random.default_rng(3)
    for rep in range(3):
        h2 = dict(host)
        for k in ("q", "r", "c"):
            h2[k] = rng.standard_normal(host[k].shape)
            inp[k] = lqr.engine.pack(h2[k])
        lqr.solve(inp, out)
        gpu = lqr.unpack_output(out)
        ref = pyoracle.lqr_factor_solve(s, h2)
        assert_lqr_parity(gpu, ref, REL_TOL)


@pytest.mark.parametrize("n,m,T", [(12, 4, 9), (5, 2, 7), (3, 2, 7), (16, 4, 6), (32, 8, 5), (64, 24, 3)])
def test_problem_major_entry_points(n, m, T):
    # sipoc_lqr_*_pm: the same three calls with problem-major device inputs -- native for
    # the CTA-per-problem plans, packed first for every other path.
    batch = 21
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=11 + n)
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, batch)
    ref = pyoracle.lqr_factor_solve(s, host)
    pm = lqr.problem_major_input(host)
    out = lqr.alloc_output()
    status = lqr.factor_solve_pm(pm, out)
    assert (status[:batch].cpu().numpy() == 0).all()
    assert_lqr_parity(lqr.unpack_output(out), ref, REL_TOL)
    # factor once, re-solve with new affine terms
    out2 = lqr.alloc_output()
    status = lqr.factor_with_status_pm(pm)
    assert (status[:batch].cpu().numpy() == 0).all()
    rng = np.random.default_rng(5)
    h2 = dict(host)
    for k in ("q", "r", "c"):
        h2[k] = rng.standard_normal(host[k].shape)
    lqr.solve_pm(lqr.problem_major_input(h2), out2)
    assert_lqr_parity(lqr.unpack_output(out2), pyoracle.lqr_factor_solve(s, h2), REL_TOL)
    # a failing problem keeps its status through the problem-major path
    bad = {k: v.copy() for k, v in host.items()}
    bad["delta"][3, :n] = -1.0
    status = lqr.factor_solve_pm(lqr.problem_major_input(bad), out)
    st = status[:batch].cpu().numpy()
    assert st[3] != 0 and (np.delete(st, 3) == 0).all()


def test_variable_dimension_chain_and_tree_batches():
    # per-node dims like tests/variable_dimensions_test.cpp:266-271, tiled
    # along a longer horizon, plus a deeper tree.
    chain = Structure.chain(9, [2, 1, 3, 2, 1, 3, 2, 1, 3, 2], [1, 2, 1, 2, 1, 2, 1, 2, 1])
    tree = Structure([0, 0, 1, 1, 2, 5], [1, 2, 3, 4, 5, 6], 0, [3, 2, 4, 1, 2, 3, 2],
                     [2, 1, 1, 2, 3, 1])
    for s in (chain, tree):
        host = pg.variable_tree_batch(s, 45, seed=11)
        ref = pyoracle.lqr_factor_solve(s, host)
        assert (ref["status"] == 0).all()
        for fused in (True, False):
            gpu, lqr = gpu_lqr_factor_solve(s, host, fused=fused)
            # chains with per-stage dims: reference-order register kernels on the padded
            # chain; trees: the generic kernels
            want = "padded_to_strict_thread_n3_m2" if s is chain else "padded_to_strict_tree_n4_m4"
            assert lqr.engine.kernel_variant == want, lqr.engine.kernel_variant
            assert (gpu["status"] == 0).all()
            assert_lqr_parity(gpu, ref, 1e-11)
            assert gpu["residual"].max() < 1e-10
            # ... which perform, on the real entries, the generic kernels' operations in
            # the same order (only the compiler's FMA contraction may differ)
            gen, lg = gpu_lqr_factor_solve(s, host, fused=fused, force_generic=True)
            assert lg.engine.kernel_variant == "generic_thread_per_problem"
            assert_lqr_parity(gpu, gen, 1e-13)
    # SIPOC_FLAG_PAD_VARIABLE_DIMS: the variable-dim chain on the (6, 2) sub-warp kernels
    # through decoupled padding; trees keep their reference-order kernels.
    host = pg.variable_tree_batch(chain, 45, seed=11)
    ref = pyoracle.lqr_factor_solve(chain, host)
    for fused in (True, False):
        gpu, lqr = gpu_lqr_factor_solve(chain, host, fused=fused, pad_variable_dims=True)
        assert lqr.engine.kernel_variant == "padded_to_subwarp4_n6_m2"
        assert (gpu["status"] == 0).all()
        assert_lqr_parity(gpu, ref, REL_TOL)
        assert gpu["residual"].max() < 1e-10
    bad = {k: v.copy() for k, v in host.items()}
    bad["delta"][7, 2] = -1.0  # node 1 of problem 7 (node 0 has two states)
    gpu, lqr = gpu_lqr_factor_solve(chain, bad, pad_variable_dims=True)
    assert gpu["status"][7] == 1 and (np.delete(gpu["status"], 7) == 0).all()
    gpu, lqr = gpu_lqr_factor_solve(tree, pg.variable_tree_batch(tree, 5, seed=1),
                                    pad_variable_dims=True)
    assert lqr.engine.kernel_variant == "padded_to_strict_tree_n4_m4"  # the flag is for chains


@pytest.mark.parametrize("shape", ["heterogeneous_chain", "shallow_wide_tree", "binary_tree"])
@pytest.mark.parametrize("num_edges,base_n", [(31, 4), (63, 4), (31, 8)])
def test_reference_variable_benchmark_shapes(shape, num_edges, base_n):
    # The VariableLQRProblem grid of lqr_benchmark.cpp:209-310, :547-555: per-node dims
    # base_n + (node % 3) - 1, per-edge controls 2 + (edge % 3) - 1, on a chain, a star
    # (every edge out of the root) and a binary tree.
    T = num_edges
    sd = [max(1, base_n + (i % 3) - 1) for i in range(T + 1)]
    cd = [max(1, 2 + (e % 3) - 1) for e in range(T)]
    children = list(range(1, T + 1))
    parents = {"heterogeneous_chain": list(range(T)), "shallow_wide_tree": [0] * T,
               "binary_tree": [(c - 1) // 2 for c in children]}[shape]
    s = pyoracle.Structure(parents, children, 0, sd, cd)
    host = pg.variable_tree_batch(s, 33, seed=17 + num_edges + base_n)
    ref = pyoracle.lqr_factor_solve(s, host)
    assert (ref["status"] == 0).all()
    for fused in (True, False):
        gpu, lqr = gpu_lqr_factor_solve(s, host, fused=fused)
        assert (gpu["status"] == 0).all()
        assert_lqr_parity(gpu, ref, REL_TOL)
        assert gpu["residual"].max() < 1e-9
    if shape == "heterogeneous_chain" and base_n == 8:
        # dims up to (9, 3): too large for the reference-order shapes (generic kernels above);
        # SIPOC_FLAG_PAD_VARIABLE_DIMS pads the chain to the (12, 4) sub-warp kernels
        gpu, lqr = gpu_lqr_factor_solve(s, host, pad_variable_dims=True)
        assert lqr.engine.kernel_variant == "padded_to_subwarp4_n12_m4"
        assert (gpu["status"] == 0).all()
        assert_lqr_parity(gpu, ref, REL_TOL)
        assert gpu["residual"].max() < 1e-9


def test_host_buffer_entry_points():
    # the reference-facing call: host arrays in, host arrays out
    n, m, T, batch = 6, 2, 12, 21
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=5, dense_M=True)
    ref = pyoracle.lqr_factor_solve(s, host)
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, batch)
    got = lqr.factor_solve_host(host)
    assert (got["status"] == 0).all()
    assert_lqr_parity(got, ref, REL_TOL)
    # factor-once / solve-many on host buffers
    st = lqr.factor_host(host)
    assert (st == 0).all()
    h2 = dict(host)
    h2["q"] = host["q"] * 0.5 + 1.0
    got2 = lqr.solve_host(h2)
    ref2 = pyoracle.lqr_factor_solve(s, h2)
    assert_lqr_parity(got2, ref2, REL_TOL)
    assert lqr.engine.launch_count > 0


def test_pack_unpack_round_trip_and_padding():
    s = Structure.chain(3, 2, 1)
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, 45)
    assert lqr.engine.batch_stride == 64
    a = np.random.default_rng(0).standard_normal((45, 77))
    dev = lqr.engine.pack(a)
    assert tuple(dev.shape) == (77, 64)
    assert np.array_equal(dev[:, :45].cpu().numpy(), a.T)
    assert (dev[:, 45:].cpu().numpy() == 0).all()
    assert np.array_equal(lqr.engine.unpack(dev, 77), a)


def test_benchmark_generator_distribution():
    # lqr_benchmark.cpp:61-96: structure of the generated data
    n, m, T, batch = 6, 2, 9, 64
    s = Structure.chain(T, n, m)
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, batch)
    inp = lqr.generate_benchmark(seed=1234, problem_offset=0)
    sz = lqr.engine.lqr_sizes
    h = {k: lqr.engine.unpack(inp[k], sz[k]) for k in sz if k in inp}
    Q = h["Q"].reshape(batch, T + 1, n, n)
    R = h["R"].reshape(batch, T, m, m)
    A = h["A"].reshape(batch, T, n, n)
    assert np.array_equal(Q, np.swapaxes(Q, -1, -2)) and np.array_equal(R, np.swapaxes(R, -1, -2))
    assert np.linalg.eigvalsh(Q).min() > 0 and np.linalg.eigvalsh(R).min() > 1.0
    assert (h["M"] == 0).all()
    assert h["delta"].min() >= 1e-3 and h["delta"].max() <= 0.101
    assert abs(h["q"].mean()) < 0.1 and abs(h["q"].std() - 1.0) < 0.1
    offdiag = A[:, :, ~np.eye(n, dtype=bool)]
    assert abs(offdiag.std() - 0.05) < 0.01 and abs(np.diagonal(A, axis1=2, axis2=3).mean() - 1) < 0.02
    # shards regenerate their own slice: offset 32 of the same seed == tail
    lqr2 = LQR(dims, topo, 32)
    inp2 = lqr2.generate_benchmark(seed=1234, problem_offset=32)
    assert np.array_equal(lqr2.engine.unpack(inp2["A"], sz["A"]), h["A"][32:])
    # and the generated problems solve to the oracle's answer
    out = lqr.alloc_output()
    st = lqr.factor_solve(inp, out)
    assert (st[:batch].cpu().numpy() == 0).all()
    ref = pyoracle.lqr_factor_solve(s, h)
    assert_lqr_parity(lqr.unpack_output(out), ref, REL_TOL)


@pytest.mark.parametrize("n,m,T,batch", [(16, 4, 24, 6000), (64, 24, 8, 600), (12, 4, 20, 20000)])
def test_specialised_kernels_at_scale_are_deterministic(n, m, T, batch):
    """Size-independent checks with every SM busy: no failed problem, every KKT residual at
    rounding level, and two runs bit-identical (a latent shared-memory race in the CTA
    kernel showed up only at full occupancy, as run-to-run differences)."""
    import torch

    dims, topo = Dimensions.uniform(T, n, m), Topology.chain(T)
    lqr = LQR(dims, topo, batch)
    inp = lqr.generate_benchmark(seed=77)
    outs = []
    for _ in range(2):
        out = lqr.alloc_output()
        status = lqr.factor_solve(inp, out)
        norms, stats = lqr.residual(inp, out, status)
        torch.cuda.synchronize()
        assert int((status[:batch] != 0).sum().item()) == 0
        assert float(stats[1].item()) < 1e-9
        outs.append(out)
    for k in ("x", "u", "y"):
        assert torch.equal(outs[0][k], outs[1][k]), k


# --- CUDA graphs over the device entry points (sipoc_graph_*) ----------------------
@pytest.mark.parametrize("n,m,T", [(4, 1, 20), (12, 4, 10), (5, 2, 7), (64, 24, 3)])
def test_lqr_factor_solve_replays_from_a_graph(n, m, T):
    """Every kernel family (thread, sub-warp, generic, CTA per problem) only enqueues on the
    caller's stream: factor + solve recorded once replays bit-exactly on new inputs."""
    import ctypes

    import torch
    from sip_optimal_control_b200._capi import lib

    batch = 37
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=11)
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, batch)
    e = lqr.engine
    inp, out = lqr.pack_input(host), lqr.alloc_output()
    status = e.empty_int()
    side = torch.cuda.Stream()
    torch.cuda.synchronize()  # the packs above ran on torch's current stream
    lqr.factor_solve(inp, out, status=status, stream=side)  # eager pass: sizes the workspaces
    side.synchronize()
    eager = lqr.unpack_output(out)
    g = ctypes.c_void_p()
    # misuse is reported, not fatal
    assert lib.sipoc_graph_end(e._handle, int(side.cuda_stream), ctypes.byref(g)) == 1
    lib.sipoc_profile_enable(e._handle, 1)
    assert lib.sipoc_graph_begin(e._handle, int(side.cuda_stream)) == 1
    lib.sipoc_profile_enable(e._handle, 0)
    e._check(lib.sipoc_graph_begin(e._handle, int(side.cuda_stream)))
    assert lib.sipoc_graph_begin(e._handle, int(side.cuda_stream)) == 1
    lqr.factor_solve(inp, out, status=status, stream=side)
    e._check(lib.sipoc_graph_end(e._handle, int(side.cuda_stream), ctypes.byref(g)))
    kernels = lib.sipoc_graph_kernel_count(g)
    assert kernels >= 1
    for k in out:
        out[k].zero_()
    torch.cuda.synchronize()
    before = e.launch_count
    e._check(lib.sipoc_graph_launch(e._handle, g, int(side.cuda_stream)))
    side.synchronize()
    assert e.launch_count == before + kernels
    replay = lqr.unpack_output(out)
    for k in ("x", "u", "y"):
        assert np.array_equal(replay[k], eager[k]), k
    # new right-hand sides written in place
    _, host2 = pg.lqr_benchmark_batch(n, m, T, batch, seed=12)
    for k in ("q", "r", "c"):
        inp[k].copy_(e.pack(host2[k]))
    torch.cuda.synchronize()
    e._check(lib.sipoc_graph_launch(e._handle, g, int(side.cuda_stream)))
    side.synchronize()
    h3 = dict(host)
    h3.update({k: host2[k] for k in ("q", "r", "c")})
    assert_lqr_parity(lqr.unpack_output(out), pyoracle.lqr_factor_solve(s, h3), REL_TOL)
    lib.sipoc_graph_destroy(g)


@pytest.mark.parametrize("n,m,T,dense_M", [(12, 4, 10, False), (6, 2, 7, True), (4, 1, 20, False)])
def test_packed_symmetric_host_entry(n, m, T, dense_M):
    # sipoc_lqr_factor_solve_host_packed: Q, R as packed lower triangles, M optional (NULL = 0);
    # the same kernels on the same numbers as the dense host entry, so bit for bit the same
    batch = 37
    s, host = pg.lqr_benchmark_batch(n, m, T, batch, seed=n + T, dense_M=dense_M)
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, batch)
    dense = lqr.factor_solve_host(host)
    packed_in = dict(host)
    packed_in["Q"] = LQR.pack_symmetric(host["Q"], n)
    packed_in["R"] = LQR.pack_symmetric(host["R"], m)
    if not dense_M:
        packed_in["M"] = None
    assert packed_in["Q"].shape[1] == (T + 1) * n * (n + 1) // 2
    got = lqr.factor_solve_host_packed(packed_in)
    assert (got["status"] == 0).all()
    for k in ("x", "u", "y"):
        assert np.array_equal(got[k], dense[k]), k
    ref = pyoracle.lqr_factor_solve(s, host)
    assert_lqr_parity({**got, "status": got["status"]}, ref, REL_TOL)
