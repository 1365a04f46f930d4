"""Committed golden vectors (tests/golden/lqr_golden.npz, made by make_golden.py):
the oracle must keep reproducing them, and the CUDA path must match them."""
import os

import numpy as np
import pytest

from gpu_helpers import REL_TOL, rel_err
from oracle import pyoracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lqr_golden.npz")


def _cases():
    z = np.load(GOLDEN)
    names = sorted({k.split("/")[0] for k in z.files})
    for name in names:
        s = pyoracle.Structure(z[f"{name}/parents"], z[f"{name}/children"],
                               int(z[f"{name}/root"][0]), z[f"{name}/state_dims"],
                               z[f"{name}/control_dims"])
        host = {k: z[f"{name}/in/{k}"] for k in pyoracle.LQR_INPUT_NAMES}
        out = {k: z[f"{name}/out/{k}"] for k in ("x", "u", "y", "residual")}
        yield name, s, host, out


def test_oracle_reproduces_golden_vectors(oracle):
    n = 0
    for name, s, host, gold in _cases():
        ref = oracle.lqr_factor_solve(s, host)
        assert (ref["status"] == 0).all(), name
        for k in ("x", "u", "y"):
            assert rel_err(ref[k], gold[k]).max() < 1e-12, (name, k)
        n += 1
    assert n == 4


@pytest.mark.gpu
def test_cuda_path_matches_golden_vectors():
    from gpu_helpers import gpu_lqr_factor_solve

    for name, s, host, gold in _cases():
        for fused in (True, False):
            gpu, _ = gpu_lqr_factor_solve(s, host, fused=fused)
            assert (gpu["status"] == 0).all(), name
            for k in ("x", "u", "y"):
                assert rel_err(gpu[k], gold[k]).max() < REL_TOL, (name, k, fused)
