"""The theta (Schur-variable) oracle pinned on the reference's own fixture
(tests/variable_dimensions_test.cpp:338-363: ||K sol - rhs|| < 1e-8) and on a dense solve
of the operator it defines; theta problems for the device tests come from here too."""
import numpy as np

import problem_gen as pg
import reference_fixtures as fx
from oracle import pyoracle
from oracle.pyoracle import Structure


def dense_operator(s, p, model, theta, w, r1, r2, r3, kd):
    """K column by column from the oracle's own y += K x."""
    eye = np.eye(kd)
    cols = [pyoracle.kkt_theta_apply(s, p, model, theta, w, r1, r2, r3, eye[j:j + 1])[0]
            for j in range(kd)]
    return np.stack(cols, axis=1)


def test_reference_schur_fixture():
    s, p, diag = fx.kkt_case_schur()
    sz = pyoracle.kkt_sizes(s)
    model, theta = fx.kkt_model(s), fx.kkt_theta_model(s, p, diag)
    w, r1, r2, r3, rhs = fx.kkt_regularization(sz["x_dim"] + p, sz["y_dim"], sz["z_dim"])
    out = pyoracle.kkt_theta_factor_solve(s, p, model, theta, w, r1, r2, r3, rhs)
    assert out["ok"].tolist() == [1]
    prod = pyoracle.kkt_theta_apply(s, p, model, theta, w, r1, r2, r3, out["sol"])
    assert np.linalg.norm(prod - rhs) < 1e-8           # the reference's own bar
    K = dense_operator(s, p, model, theta, w, r1, r2, r3, rhs.shape[1])
    assert np.abs(K - K.T).max() < 1e-15               # the operator is symmetric
    dense = np.linalg.solve(K, rhs[0])
    assert np.abs(out["sol"][0] - dense).max() / np.abs(dense).max() < 1e-12


def test_random_theta_problems_match_dense_solve():
    s = Structure.chain(5, [3, 2, 4, 3, 2, 3], [2, 1, 2, 1, 2], node_c=[1, 0, 2, 0, 1, 1],
                        node_g=[0, 2, 1, 0, 0, 2], edge_c=[1, 0, 2, 1, 0], edge_g=[2, 1, 0, 1, 2])
    p, batch = 3, 4
    model, theta, w, r1, r2, r3, rhs = pg.newton_kkt_theta_batch(s, p, batch, seed=3, r2_max=1e2)
    out = pyoracle.kkt_theta_factor_solve(s, p, model, theta, w, r1, r2, r3, rhs)
    assert out["ok"].all()
    for b in range(batch):
        one = lambda d: {k: v[b:b + 1] for k, v in d.items()}
        K = dense_operator(s, p, one(model), one(theta), w[b:b + 1], r1[b:b + 1], r2[b:b + 1],
                           r3[b:b + 1], rhs.shape[1])
        dense = np.linalg.solve(K, rhs[b])
        assert np.abs(out["sol"][b] - dense).max() / np.abs(dense).max() < 1e-10


def test_indefinite_schur_complement_fails_the_factor():
    s, p, _ = fx.kkt_case_schur()
    sz = pyoracle.kkt_sizes(s)
    model, theta = fx.kkt_model(s), fx.kkt_theta_model(s, p, -50.0)  # S = sum H_tt + ... < 0
    w, r1, r2, r3, rhs = fx.kkt_regularization(sz["x_dim"] + p, sz["y_dim"], sz["z_dim"])
    out = pyoracle.kkt_theta_factor_solve(s, p, model, theta, w, r1, r2, r3, rhs)
    assert out["ok"].tolist() == [0]
