"""Pin the LQR oracle on the reference's own test fixtures (tests/lqr_test.cpp).

Each test restates one gtest case with the reference's acceptance bar.
"""
import numpy as np
import pytest

import reference_fixtures as fx
from oracle.pyoracle import Structure

SUCCESS, INVALID_DELTA, F_FAIL, G_FAIL, INVALID_TOPOLOGY = 0, 1, 2, 3, 4


def _solve(oracle, s, p, **kw):
    return oracle.lqr_factor_solve(s, fx.pack_problem(**p), **kw)


def test_factor_reports_success(oracle):  # lqr_test.cpp:188-192
    s, p = fx.identity_chain(2, 1, 2)
    assert _solve(oracle, s, p, solve=False)["status"][0] == SUCCESS


def test_factor_reports_invalid_delta(oracle):  # lqr_test.cpp:206-211
    s, p = fx.identity_chain(2, 1, 2)
    p["delta"][2][0] = 0.0
    assert _solve(oracle, s, p, solve=False)["status"][0] == INVALID_DELTA


def test_factor_reports_f_factorization_failure(oracle):  # lqr_test.cpp:213-219
    s, p = fx.identity_chain(1, 1, 1)
    p["Q"][1][0, 0] = -2.0
    p["delta"][1][0] = 1.0
    assert _solve(oracle, s, p, solve=False)["status"][0] == F_FAIL


def test_factor_reports_g_factorization_failure(oracle):  # lqr_test.cpp:221-227
    s, p = fx.identity_chain(1, 1, 1)
    p["Q"][1][0, 0] = 0.0
    p["R"][0][0, 0] = -1.0
    assert _solve(oracle, s, p, solve=False)["status"][0] == G_FAIL


def test_solves_nonuniform_diagonal_delta_problem(oracle):  # lqr_test.cpp:229-263
    s, p = fx.nonuniform_delta_chain()
    out = _solve(oracle, s, p)
    assert out["status"][0] == SUCCESS
    assert out["residual"][0] < 1e-12


def test_solves_branching_tree_problem(oracle):  # lqr_test.cpp:411-429
    s, p = fx.branch_tree()
    out = _solve(oracle, s, p)
    assert out["status"][0] == SUCCESS
    assert out["residual"][0] < 1e-12


def test_factor_and_solve_are_repeatable(oracle):  # lqr_test.cpp:431-450
    s, p = fx.branch_tree()
    a = _solve(oracle, s, p, repeats=2)
    b = _solve(oracle, s, p)
    for k in ("x", "u", "y"):
        assert np.array_equal(a[k], b[k])


def test_rejects_invalid_tree_topology(oracle):  # lqr_test.cpp:452-464
    s = Structure([0, 0], [1, 1], 0, [2, 2, 2], [1, 1])
    assert oracle.compile_topology(s)[0] == INVALID_TOPOLOGY


def test_solves_variable_dimension_branching_tree(oracle):  # lqr_test.cpp:641-659
    s, p = fx.variable_dim_branch_tree()
    out = _solve(oracle, s, p)
    assert out["status"][0] == SUCCESS
    assert out["residual"][0] < 1e-12


def test_compiles_multi_child_preorder_and_postorder(oracle):  # lqr_test.cpp:931-953
    s, _ = fx.five_node_tree()
    st, child_offsets, child_edges, pre, post = oracle.compile_topology(s)
    assert st == SUCCESS
    assert child_offsets.tolist() == [0, 2, 4, 4, 4, 4]
    assert child_edges.tolist() == [0, 1, 2, 3]
    assert pre.tolist() == [0, 1, 3, 4, 2]
    assert post.tolist() == [2, 4, 3, 1, 0]


def test_rejects_disconnected_tree(oracle):  # lqr_test.cpp:955-967
    s, _ = fx.five_node_tree(parents=(0, 0, 1, 4), children=(1, 2, 3, 3))
    assert oracle.compile_topology(s)[0] == INVALID_TOPOLOGY


def test_rejects_cycle(oracle):  # lqr_test.cpp:969-980
    s, _ = fx.five_node_tree(parents=(4, 0, 1, 1))
    assert oracle.compile_topology(s)[0] == INVALID_TOPOLOGY


def _is_approx(a, b, prec):
    # Eigen isApprox: ||a - b||^2 <= prec^2 * min(||a||^2, ||b||^2)
    return np.sum((a - b) ** 2) <= prec * prec * min(np.sum(a * a), np.sum(b * b))


def test_matches_dense_kkt_on_variable_dimension_tree(oracle):  # lqr_test.cpp:982-1013
    s, p = fx.five_node_tree()
    out = _solve(oracle, s, p)
    assert out["status"][0] == SUCCESS
    xs, us, ys = fx.dense_kkt_solve(s, p)
    no = np.concatenate([[0], np.cumsum(s.state_dims)])
    mo = np.concatenate([[0], np.cumsum(s.control_dims)])
    for node in range(5):
        assert _is_approx(out["x"][0, no[node]:no[node + 1]], xs[node], 1e-10)
        assert _is_approx(out["y"][0, no[node]:no[node + 1]], ys[node], 1e-10)
    for e in range(4):
        assert _is_approx(out["u"][0, mo[e]:mo[e + 1]], us[e], 1e-10)
    assert out["residual"][0] < 1e-12


@pytest.mark.parametrize("builder", [fx.nonuniform_delta_chain, fx.branch_tree,
                                     fx.variable_dim_branch_tree])
def test_dense_kkt_agrees_on_every_fixture(oracle, builder):
    s, p = builder()
    out = _solve(oracle, s, p)
    xs, us, ys = fx.dense_kkt_solve(s, p)
    assert np.allclose(out["x"][0], np.concatenate(xs), rtol=0, atol=1e-12)
    assert np.allclose(out["u"][0], np.concatenate(us), rtol=0, atol=1e-12)
    assert np.allclose(out["y"][0], np.concatenate(ys), rtol=0, atol=1e-12)


def test_first_failure_in_postorder_wins(oracle):
    # lqr.cpp:696-700,722-727: at a node the child edge's G test precedes the
    # node's own delta / F tests; nodes are visited leaf-to-root.
    s, p = fx.identity_chain(2, 1, 3)
    p["delta"][0][1] = -1.0          # root: would be INVALID_DELTA (visited last)
    p["R"][2][0, 0] = -50.0          # edge 2 (visited first): G failure
    assert _solve(oracle, s, p, solve=False)["status"][0] == G_FAIL
    s, p = fx.identity_chain(2, 1, 3)
    p["delta"][3][0] = 0.0           # terminal node delta, visited before edge 2's G
    p["R"][2][0, 0] = -50.0
    assert _solve(oracle, s, p, solve=False)["status"][0] == INVALID_DELTA


def test_batch_is_elementwise_independent(oracle):
    rng = np.random.default_rng(0)
    s, p = fx.nonuniform_delta_chain()
    one = fx.pack_problem(**p)
    batch = {k: np.repeat(v, 5, axis=0) for k, v in one.items()}
    batch["q"] = batch["q"] + rng.standard_normal(batch["q"].shape)
    out = oracle.lqr_factor_solve(s, batch, nthreads=3)
    for b in range(5):
        single = oracle.lqr_factor_solve(s, {k: v[b:b + 1] for k, v in batch.items()})
        assert np.array_equal(single["x"][0], out["x"][b])
    assert np.all(out["residual"] < 1e-12)
