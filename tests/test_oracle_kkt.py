"""Pin the Newton-KKT oracle on tests/variable_dimensions_test.cpp (theta == 0)."""
import numpy as np
import pytest

import reference_fixtures as fx


def _run(oracle, s):
    sz = oracle.kkt_sizes(s)
    model = fx.kkt_model(s)
    w, r1, r2, r3, rhs = fx.kkt_regularization(sz["x_dim"], sz["y_dim"], sz["z_dim"])
    out = oracle.kkt_factor_solve(s, model, w, r1, r2, r3, rhs)
    return sz, model, (w, r1, r2, r3, rhs), out


def test_wire_format_offsets(oracle):
    # types.cpp:24-64 on the dims of variable_dimensions_test.cpp:266-271.
    s = fx.kkt_case_chain()
    sz = oracle.kkt_sizes(s)
    assert (sz["x_dim"], sz["y_dim"], sz["z_dim"], sz["kkt_dim"]) == (9, 12, 6, 27)
    o = oracle.kkt_offsets(s)
    assert o["x_state"].tolist() == [0, 3, 6]
    assert o["x_control"].tolist() == [2, 4]
    assert o["y_dyn"].tolist() == [0, 3, 4]
    assert o["y_node_c"].tolist() == [2, 4, 7]
    assert o["y_edge_c"].tolist() == [9, 10]
    assert o["z_node"].tolist() == [0, 0, 2]
    assert o["z_edge"].tolist() == [3, 5]


@pytest.mark.parametrize("case", [fx.kkt_case_chain, fx.kkt_case_siblings,
                                  fx.kkt_case_zero_dim_root])
def test_callback_provider_solves(oracle, case):
    # variable_dimensions_test.cpp:265-336 through expect_kkt_solve (:135-181):
    # factor must succeed and ||K sol - rhs||_2 < 1e-9.
    s = case()
    _, _, _, out = _run(oracle, s)
    assert out["ok"][0] == 1
    assert out["residual"][0] < 1e-9


@pytest.mark.parametrize("case", [fx.kkt_case_chain, fx.kkt_case_siblings])
def test_matches_dense_solve_of_the_kkt_operator(oracle, case):
    # Assemble K column by column from add_Kx_to_y and solve densely.
    s = case()
    sz, model, (w, r1, r2, r3, rhs), out = _run(oracle, s)
    kd = sz["kkt_dim"]
    eye = np.eye(kd)
    rep = lambda a: np.repeat(a, kd, axis=0)
    Kmat = oracle.kkt_apply(s, {k: rep(v) for k, v in model.items()}, rep(w), rep(r1),
                            rep(r2), rep(r3), eye).T
    assert np.allclose(Kmat, Kmat.T, atol=1e-14)
    dense = np.linalg.solve(Kmat, rhs[0])
    assert np.linalg.norm(out["sol"][0] - dense) <= 1e-12 * np.linalg.norm(dense)


def test_factor_rejects_nonpositive_regularization(oracle):
    # helpers.cpp:251-295: r2 <= 0 or w + r3 <= 0 -> factor returns false.
    s = fx.kkt_case_chain()
    sz = oracle.kkt_sizes(s)
    model = fx.kkt_model(s)
    w, r1, r2, r3, rhs = fx.kkt_regularization(sz["x_dim"], sz["y_dim"], sz["z_dim"])
    bad = r2.copy()
    bad[0, 5] = 0.0
    assert oracle.kkt_factor_solve(s, model, w, r1, bad, r3, rhs)["ok"][0] == 0
    bad = r3.copy()
    bad[0, 2] = -1.3
    assert oracle.kkt_factor_solve(s, model, w, r1, r2, bad, rhs)["ok"][0] == 0
    assert oracle.kkt_factor_solve(s, model, w, r1, r2, r3, rhs)["ok"][0] == 1
