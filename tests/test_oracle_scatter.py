"""Oracle restatement of the model-callback scatter (sip_optimal_control.cpp:44-123) against
an independent numpy statement of the same lines, on the reference's branched
variable-dimension structures (variable_dimensions_test.cpp:292-336, 338-363)."""
import numpy as np
import pytest

import reference_fixtures as fx
from oracle import pyoracle


def random_values(s, p, batch, seed):
    rng = np.random.default_rng(seed)
    sz = pyoracle.model_value_sizes(s, p)
    vals = {k: rng.standard_normal((batch, sz[k])) for k in pyoracle.MODEL_VALUE_NAMES}
    ksz = pyoracle.kkt_sizes(s)
    x = rng.standard_normal((batch, ksz["x_dim"] + p))
    x0 = rng.standard_normal((batch, s.state_dims[s.root]))
    return vals, x, x0


def numpy_scatter(s, p, vals, x, x0):
    """The reference's loops, one problem at a time, straight from the offsets table."""
    o, ksz = pyoracle.kkt_offsets(s), pyoracle.kkt_sizes(s)
    E, N = s.num_edges, s.num_edges + 1
    n, m = list(s.state_dims), list(s.control_dims)
    batch = x.shape[0]
    f = np.zeros(batch)
    grad = np.zeros((batch, ksz["x_dim"] + p))
    c = np.zeros((batch, ksz["y_dim"]))
    g = np.zeros((batch, ksz["z_dim"]))
    for i in range(batch):
        cur = {k: 0 for k in pyoracle.MODEL_VALUE_NAMES}

        def take(k, count):
            out = vals[k][i, cur[k]:cur[k] + count]
            cur[k] += count
            return out
        acc = 0.0
        for node in range(N):
            acc += vals["node_f"][i, node]
        for e in range(E):
            acc += vals["edge_f"][i, e]
        f[i] = acc
        for node in range(N):
            grad[i, o["x_state"][node]:o["x_state"][node] + n[node]] += take("node_df_dx", n[node])
            grad[i, ksz["x_dim"]:] += take("node_df_dtheta", p)
            c[i, o["y_node_c"][node]:o["y_node_c"][node] + s.node_c[node]] = take("node_c", s.node_c[node])
            g[i, o["z_node"][node]:o["z_node"][node] + s.node_g[node]] = take("node_g", s.node_g[node])
        for e in range(E):
            par, chi = s.parents[e], s.children[e]
            grad[i, o["x_state"][par]:o["x_state"][par] + n[par]] += take("edge_df_dx", n[par])
            grad[i, o["x_control"][e]:o["x_control"][e] + m[e]] += take("edge_df_du", m[e])
            grad[i, ksz["x_dim"]:] += take("edge_df_dtheta", p)
            c[i, o["y_dyn"][chi]:o["y_dyn"][chi] + n[chi]] = take("edge_dyn_res", n[chi])
            c[i, o["y_edge_c"][e]:o["y_edge_c"][e] + s.edge_c[e]] = take("edge_c", s.edge_c[e])
            g[i, o["z_edge"][e]:o["z_edge"][e] + s.edge_g[e]] = take("edge_g", s.edge_g[e])
        r = s.root
        c[i, o["y_dyn"][r]:o["y_dyn"][r] + n[r]] = x0[i] - x[i, o["x_state"][r]:o["x_state"][r] + n[r]]
    return dict(f=f, gradient_f=grad, c=c, g=g)


CASES = {"chain": (fx.kkt_case_chain, 0), "siblings": (fx.kkt_case_siblings, 0),
         "zero_dim_root": (fx.kkt_case_zero_dim_root, 0),
         "schur": (lambda: fx.kkt_case_schur()[0], 2)}


@pytest.mark.parametrize("name", sorted(CASES))
def test_scatter_matches_numpy_statement(name):
    make, p = CASES[name]
    s = make()
    vals, x, x0 = random_values(s, p, 5, seed=len(name))
    got = pyoracle.model_scatter(s, vals, x, x0, p)
    want = numpy_scatter(s, p, vals, x, x0)
    for k in ("f", "gradient_f", "c", "g"):
        assert np.array_equal(got[k], want[k]), k


def test_scatter_without_new_x_computes_f_only():
    s = fx.kkt_case_siblings()
    vals, x, x0 = random_values(s, 0, 3, seed=1)
    got = pyoracle.model_scatter(s, vals, x, x0, 0, new_x=False)
    assert np.array_equal(got["f"], numpy_scatter(s, 0, vals, x, x0)["f"])
    assert not got["gradient_f"].any() and not got["c"].any() and not got["g"].any()


def test_every_output_entry_is_written_once():
    # the y and z layouts are covered exactly: no entry is left at its fill value
    s = fx.kkt_case_siblings()
    sz = pyoracle.model_value_sizes(s, 0)
    vals = {k: np.full((1, sz[k]), 1.0 + i) for i, k in enumerate(pyoracle.MODEL_VALUE_NAMES)}
    ksz = pyoracle.kkt_sizes(s)
    x = np.zeros((1, ksz["x_dim"]))
    x0 = np.full((1, s.state_dims[s.root]), 99.0)
    got = pyoracle.model_scatter(s, vals, x, x0)
    assert (got["c"] != 0).all() and (got["g"] != 0).all() and (got["gradient_f"] != 0).all()
