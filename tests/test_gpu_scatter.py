"""GPU parity of the batched model-callback scatter (sipoc_model_scatter; the model_callback
lambda of sip_optimal_control.cpp:13-127) against the CPU oracle: pure data movement and sums
in the reference's order, so the comparison is bit for bit."""
import numpy as np
import pytest

import reference_fixtures as fx
from oracle import pyoracle
from oracle.pyoracle import Structure
from sip_optimal_control_b200 import CallbackProvider, Dimensions, Topology
from test_oracle_scatter import random_values

pytestmark = pytest.mark.gpu


def _provider(s, p, batch):
    topo = Topology(s.num_edges, s.root, s.parents, s.children)
    dims = Dimensions(p, s.state_dims, s.control_dims, s.node_c, s.node_g, s.edge_c, s.edge_g)
    return CallbackProvider(dims, topo, batch)


def _quadrotor_kkt_chain(T=50, n=12, m=4):
    c, g = max(1, n // 2), max(1, 2 * m)  # newton_kkt_benchmark.cpp:63-64, 79-80
    return Structure.chain(T, n, m, node_c=[0] * T + [c], node_g=[0] * T + [g], edge_c=[c] * T,
                           edge_g=[g] * T)


CASES = {
    "chain": (fx.kkt_case_chain, 0, 37), "siblings": (fx.kkt_case_siblings, 0, 19),
    "zero_dim_root": (fx.kkt_case_zero_dim_root, 0, 5),
    "schur": (lambda: fx.kkt_case_schur()[0], 2, 33),
    "quadrotor_kkt": (_quadrotor_kkt_chain, 0, 130),
    "quadrotor_kkt_theta": (lambda: _quadrotor_kkt_chain(10), 8, 64),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_scatter_bit_exact(name):
    make, p, batch = CASES[name]
    s = make()
    vals, x, x0 = random_values(s, p, batch, seed=batch)
    ref = pyoracle.model_scatter(s, vals, x, x0, p)
    cp = _provider(s, p, batch)
    e = cp.engine
    assert cp.model_value_sizes == pyoracle.model_value_sizes(s, p)
    dv = {k: e.pack(v) for k, v in vals.items()}
    out = cp.model_callback_scatter(dv, e.pack(x), e.pack(x0))
    sz = cp.sizes
    assert np.array_equal(e.unpack(out["f"], 1)[:, 0], ref["f"])
    for k, dim in (("gradient_f", sz["x_dim"]), ("c", sz["y_dim"]), ("g", sz["z_dim"])):
        assert np.array_equal(e.unpack(out[k], dim), ref[k]), k


@pytest.mark.parametrize("name", ["siblings", "schur", "quadrotor_kkt"])
def test_host_scatter_bit_exact(name):
    make, p, batch = CASES[name]
    s = make()
    vals, x, x0 = random_values(s, p, batch, seed=3)
    ref = pyoracle.model_scatter(s, vals, x, x0, p)
    cp = _provider(s, p, batch)
    got = cp.model_callback_scatter_host(vals, x, x0)
    for k in ("f", "gradient_f", "c", "g"):
        assert np.array_equal(got[k], ref[k]), k
    # new_x == false: the objective alone (sip_optimal_control.cpp:52)
    only_f = cp.model_callback_scatter_host(vals, x, x0, new_x=False)
    assert np.array_equal(only_f["f"], ref["f"])
    assert not only_f["gradient_f"].any() and not only_f["c"].any()
