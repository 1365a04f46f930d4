"""Randomised structures against the CPU oracle: random trees, per-node state / per-edge control
dimensions, random constraint dimensions and ragged batch
sizes, on whatever kernels the engine dispatches by itself and on the generic ones.  Seeds are
fixed, so a failure reproduces."""
import numpy as np
import pytest

import problem_gen as pg
from gpu_helpers import REL_TOL, assert_lqr_parity, gpu_lqr_factor_solve, rel_err
from oracle import pyoracle
from oracle.pyoracle import Structure
from test_gpu_kkt import _gpu_kkt

pytestmark = pytest.mark.gpu


def random_structure(rng, max_nodes=12, max_n=5, max_m=3, constraints=False):
    N = int(rng.integers(1, max_nodes + 1))
    # node i > 0 hangs off a random earlier node; edges in a random order, nodes relabelled
    perm = rng.permutation(N)
    parents, children = [], []
    for i in range(1, N):
        parents.append(int(perm[rng.integers(0, i)]))
        children.append(int(perm[i]))
    order = rng.permutation(N - 1) if N > 1 else np.array([], int)
    parents = [parents[k] for k in order]
    children = [children[k] for k in order]
    sd = [int(v) for v in rng.integers(1, max_n + 1, size=N)]
    cd = [int(v) for v in rng.integers(1, max_m + 1, size=max(N - 1, 0))]
    kw = {}
    if constraints:
        kw = dict(node_c=[int(v) for v in rng.integers(0, 3, size=N)],
                  node_g=[int(v) for v in rng.integers(0, 3, size=N)],
                  edge_c=[int(v) for v in rng.integers(0, 3, size=max(N - 1, 0))],
                  edge_g=[int(v) for v in rng.integers(0, 3, size=max(N - 1, 0))])
    return Structure(parents, children, int(perm[0]), sd, cd, **kw)


@pytest.mark.parametrize("seed", range(16))
def test_random_trees_lqr(seed):
    rng = np.random.default_rng(1000 + seed)
    s = random_structure(rng)
    batch = int(rng.integers(1, 71))
    host = pg.variable_tree_batch(s, batch, seed=seed)
    ref = pyoracle.lqr_factor_solve(s, host)
    assert (ref["status"] == 0).all()
    for force_generic in (False, True):
        gpu, lqr = gpu_lqr_factor_solve(s, host, force_generic=force_generic)
        assert (gpu["status"] == 0).all(), lqr.engine.kernel_variant
        assert_lqr_parity(gpu, ref, REL_TOL)


@pytest.mark.parametrize("seed", range(10))
def test_random_trees_newton_kkt(seed):
    rng = np.random.default_rng(2000 + seed)
    s = random_structure(rng, max_nodes=8, max_n=4, max_m=2, constraints=True)
    batch = int(rng.integers(1, 50))
    model, w, r1, r2, r3, rhs = pg.newton_kkt_batch(s, batch, seed=seed, r2_max=1e6)
    ref = pyoracle.kkt_factor_solve(s, model, w, r1, r2, r3, rhs)
    for force_generic in (False, True):
        gpu, cp, _ = _gpu_kkt(s, model, w, r1, r2, r3, rhs, force_generic=force_generic)
        assert (gpu["ok"] == ref["ok"]).all(), cp.engine.kernel_variant
        good = ref["ok"] == 1
        assert good.any()
        assert rel_err(gpu["sol"][good], ref["sol"][good]).max() < REL_TOL, cp.engine.kernel_variant
