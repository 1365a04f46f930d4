"""Generates tests/golden/lqr_golden.npz: seeded inputs and the ORACLE's outputs for a few
small problems, committed so that (i) the oracle is pinned against its own regressions
and (ii) the GPU box can check the CUDA path against stored vectors without trusting a
freshly built oracle.  The reference itself cannot be run here (it needs Eigen and sip,
see DESIGN.md section 4), so these are outputs of the oracle after it passed the
reference's closed-form fixtures (tests/test_oracle_lqr.py, tests/test_oracle_kkt.py).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import problem_gen as pg  # noqa: E402
from oracle import pyoracle  # noqa: E402

CASES = {
    "chain_n4_m1_T10": lambda: pg.lqr_benchmark_batch(4, 1, 10, 3, seed=101),
    "chain_n12_m4_T6": lambda: pg.lqr_benchmark_batch(12, 4, 6, 3, seed=102, dense_M=True),
    "chain_n16_m4_T3": lambda: pg.lqr_benchmark_batch(16, 4, 3, 2, seed=103),
}


def tree_case():
    s = pyoracle.Structure([0, 0, 1, 1], [1, 2, 3, 4], 0, [3, 1, 2, 4, 2], [2, 1, 3, 1])
    return s, pg.variable_tree_batch(s, 3, seed=104)


def main():
    out = {}
    cases = dict(CASES)
    cases["tree_variable_dims"] = tree_case
    for name, make in cases.items():
        s, host = make()
        ref = pyoracle.lqr_factor_solve(s, host)
        assert (ref["status"] == 0).all()
        out[f"{name}/parents"] = s.parents
        out[f"{name}/children"] = s.children
        out[f"{name}/root"] = np.array([s.root])
        out[f"{name}/state_dims"] = s.state_dims
        out[f"{name}/control_dims"] = s.control_dims
        for k, v in host.items():
            out[f"{name}/in/{k}"] = v
        for k in ("x", "u", "y", "residual"):
            out[f"{name}/out/{k}"] = ref[k]
    np.savez_compressed(os.path.join(HERE, "lqr_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "lqr_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
