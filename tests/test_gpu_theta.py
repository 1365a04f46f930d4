"""GPU parity tests of the theta (global / Schur variable) layer of the Newton-KKT path:
CallbackProvider::factor / solve / add_Kx_to_y / add_*x_to_y with theta_dim > 0
(helpers.cpp:190-240, 372-407, 896-951, theta branches of 1019-1368), through the C ABI."""
import numpy as np
import pytest

import problem_gen as pg
import reference_fixtures as fx
from gpu_helpers import REL_TOL, rel_err
from oracle import pyoracle
from oracle.pyoracle import Structure
from sip_optimal_control_b200 import CallbackProvider, Dimensions, Topology

pytestmark = pytest.mark.gpu


def _provider(s, p, batch, **kw):
    topo = Topology(s.num_edges, s.root, s.parents, s.children)
    dims = Dimensions(p, s.state_dims, s.control_dims, s.node_c, s.node_g, s.edge_c, s.edge_g)
    return CallbackProvider(dims, topo, batch, **kw)


def _gpu_theta(s, p, model, theta, w, r1, r2, r3, rhs, **kw):
    batch = rhs.shape[0]
    cp = _provider(s, p, batch, **kw)
    e = cp.engine
    assert cp.sizes["theta_dim"] == p and cp.sizes["kkt_dim"] == rhs.shape[1]
    dm = cp.pack_model({**model, **theta})
    dw, dr1, dr2, dr3, db = (e.pack(a) for a in (w, r1, r2, r3, rhs))
    ok = cp.factor(dm, dw, dr1, dr2, dr3)
    sol = e.zeros(rhs.shape[1])
    cp.solve(dm, db, sol)
    norms, stats = cp.residual(dm, dw, dr1, dr2, dr3, sol, db, ok)
    return dict(sol=e.unpack(sol, rhs.shape[1]), ok=ok[:batch].cpu().numpy(),
                residual=norms[:batch].cpu().numpy()), cp, (dm, dw, dr1, dr2, dr3, db)


def test_reference_schur_fixture():
    # variable_dimensions_test.cpp:338-363 (SolvesBranchedSystemWithSchurVariables)
    s, p, diag = fx.kkt_case_schur()
    sz = pyoracle.kkt_sizes(s)
    rep = lambda a: np.repeat(a, 3, axis=0)
    model = {k: rep(v) for k, v in fx.kkt_model(s).items()}
    theta = {k: rep(v) for k, v in fx.kkt_theta_model(s, p, diag).items()}
    w, r1, r2, r3, rhs = (rep(a) for a in
                          fx.kkt_regularization(sz["x_dim"] + p, sz["y_dim"], sz["z_dim"]))
    ref = pyoracle.kkt_theta_factor_solve(s, p, model, theta, w, r1, r2, r3, rhs)
    gpu, cp, _ = _gpu_theta(s, p, model, theta, w, r1, r2, r3, rhs)
    assert gpu["ok"].tolist() == [1, 1, 1]
    assert gpu["residual"].max() < 1e-8              # the reference's own bar
    assert rel_err(gpu["sol"], ref["sol"]).max() < 1e-12


@pytest.mark.parametrize("n,m,T,p", [(4, 1, 16, 4), (12, 4, 10, 8), (6, 2, 12, 4), (16, 4, 5, 4)])
@pytest.mark.parametrize("force_generic", [True, False])
def test_theta_benchmark_shapes(n, m, T, p, force_generic):
    # BM_NewtonKKTTheta* shapes (newton_kkt_benchmark.cpp:253-263): theta_dim in {4, 8}
    c, g = max(1, n // 2), max(1, 2 * m)
    s = Structure.chain(T, n, m, node_c=[0] * T + [c], node_g=[0] * T + [g], edge_c=[c] * T,
                        edge_g=[g] * T)
    batch = 21
    model, theta, w, r1, r2, r3, rhs = pg.newton_kkt_theta_batch(s, p, batch, seed=n + p,
                                                                 r2_max=1e9)
    ref = pyoracle.kkt_theta_factor_solve(s, p, model, theta, w, r1, r2, r3, rhs)
    good = ref["ok"] == 1
    assert good.sum() >= batch - 2
    gpu, cp, dev = _gpu_theta(s, p, model, theta, w, r1, r2, r3, rhs, force_generic=force_generic)
    assert (gpu["ok"] == ref["ok"]).all()
    assert rel_err(gpu["sol"][good], ref["sol"][good]).max() < REL_TOL
    # the operator with its theta rows and columns
    dm, dw, dr1, dr2, dr3, _ = dev
    xv = np.random.default_rng(1).standard_normal(rhs.shape)
    y0 = np.random.default_rng(2).standard_normal(rhs.shape)
    dy = cp.engine.pack(y0)
    cp.add_Kx_to_y(dm, dw, dr1, dr2, dr3, cp.engine.pack(xv), dy)
    yref = pyoracle.kkt_theta_apply(s, p, model, theta, w, r1, r2, r3, xv, y0)
    assert rel_err(cp.engine.unpack(dy, rhs.shape[1]), yref).max() < 1e-12


def test_theta_on_a_variable_dimension_tree_and_operator_blocks():
    s = Structure([0, 0, 1, 1], [1, 2, 3, 4], 0, [3, 1, 2, 4, 2], [2, 1, 3, 1],
                  node_c=[1, 0, 2, 0, 1], node_g=[0, 2, 1, 1, 0], edge_c=[1, 0, 2, 1],
                  edge_g=[2, 1, 0, 1])
    p, batch = 3, 9
    model, theta, w, r1, r2, r3, rhs = pg.newton_kkt_theta_batch(s, p, batch, seed=5, r2_max=1e3)
    ref = pyoracle.kkt_theta_factor_solve(s, p, model, theta, w, r1, r2, r3, rhs)
    assert ref["ok"].all()
    gpu, cp, dev = _gpu_theta(s, p, model, theta, w, r1, r2, r3, rhs)
    assert gpu["ok"].all()
    assert rel_err(gpu["sol"], ref["sol"]).max() < REL_TOL
    # add_Hx / Cx / CTx / Gx / GTx against the oracle's operator with the regularization
    # off and one slot of [x | y | z] populated at a time
    e = cp.engine
    dm = dev[0]
    sz = pyoracle.kkt_sizes(s)
    xd, yd, zd = sz["x_dim"] + p, sz["y_dim"], sz["z_dim"]
    rng = np.random.default_rng(8)
    zero = lambda k: np.zeros((batch, k))
    vx, vy, vz = (rng.standard_normal((batch, k)) for k in (xd, yd, zd))

    def oracle(x_x, x_y, x_z):
        x = np.concatenate([x_x, x_y, x_z], axis=1)
        y = pyoracle.kkt_theta_apply(s, p, model, theta, zero(zd), zero(xd), zero(yd), zero(zd), x)
        return y[:, :xd], y[:, xd:xd + yd], y[:, xd + yd:]

    cases = [(cp.add_Hx_to_y, vx, xd, oracle(vx, zero(yd), zero(zd))[0]),
             (cp.add_Cx_to_y, vx, yd, oracle(vx, zero(yd), zero(zd))[1]),
             (cp.add_CTx_to_y, vy, xd, oracle(zero(xd), vy, zero(zd))[0]),
             (cp.add_Gx_to_y, vx, zd, oracle(vx, zero(yd), zero(zd))[2]),
             (cp.add_GTx_to_y, vz, xd, oracle(zero(xd), zero(yd), vz)[0])]
    for fn, vin, nout, want in cases:
        y0 = rng.standard_normal((batch, nout))
        dy = e.pack(y0)
        fn(dm, e.pack(vin), dy)
        assert np.abs(e.unpack(dy, nout) - (y0 + want)).max() < 1e-12, fn.__name__


def test_indefinite_schur_complement_is_reported_per_problem():
    s, p, diag = fx.kkt_case_schur()
    sz = pyoracle.kkt_sizes(s)
    batch = 4
    rep = lambda a: np.repeat(a, batch, axis=0)
    model = {k: rep(v) for k, v in fx.kkt_model(s).items()}
    theta = {k: rep(v) for k, v in fx.kkt_theta_model(s, p, diag).items()}
    bad = fx.kkt_theta_model(s, p, -50.0)
    for k in theta:
        theta[k][2] = bad[k][0]
    w, r1, r2, r3, rhs = (rep(a) for a in
                          fx.kkt_regularization(sz["x_dim"] + p, sz["y_dim"], sz["z_dim"]))
    ref = pyoracle.kkt_theta_factor_solve(s, p, model, theta, w, r1, r2, r3, rhs)
    gpu, _, _ = _gpu_theta(s, p, model, theta, w, r1, r2, r3, rhs)
    assert ref["ok"].tolist() == [1, 1, 0, 1]
    assert gpu["ok"].tolist() == [1, 1, 0, 1]
    good = ref["ok"] == 1
    assert rel_err(gpu["sol"][good], ref["sol"][good]).max() < 1e-12


def test_theta_host_buffer_entry_points():
    s, p, diag = fx.kkt_case_schur()
    sz = pyoracle.kkt_sizes(s)
    model, theta = fx.kkt_model(s), fx.kkt_theta_model(s, p, diag)
    w, r1, r2, r3, rhs = fx.kkt_regularization(sz["x_dim"] + p, sz["y_dim"], sz["z_dim"])
    ref = pyoracle.kkt_theta_factor_solve(s, p, model, theta, w, r1, r2, r3, rhs)
    cp = _provider(s, p, 1)
    both = {**model, **theta}
    # the operator works from an uploaded model alone, before any factor (helpers.cpp:1161-1183)
    cp.set_model_host(both)
    prod0 = cp.add_Kx_to_y_host(w, r1, r2, r3, ref["sol"])
    assert np.linalg.norm(prod0 - rhs) < 1e-8
    ok = cp.factor_host(both, w, r1, r2, r3)
    assert ok.tolist() == [1]
    sol = cp.solve_host(rhs)
    assert rel_err(sol, ref["sol"]).max() < 1e-12
