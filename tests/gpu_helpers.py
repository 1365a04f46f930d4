"""Shared helpers of the ``-m gpu`` parity tests: run the CUDA path through the
C ABI and compare it with the CPU oracle on identical inputs."""
from __future__ import annotations

import numpy as np

from oracle import pyoracle
from sip_optimal_control_b200 import LQR, CallbackProvider, Dimensions, Topology

# north_star: FP64 results within 1e-9 relative of the reference path.
REL_TOL = 1e-9


def to_structs(s: pyoracle.Structure):
    topo = Topology(s.num_edges, s.root, s.parents, s.children)
    dims = Dimensions(0, s.state_dims, s.control_dims, s.node_c, s.node_g, s.edge_c, s.edge_g)
    return dims, topo


def rel_err(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Per-problem ||a - b||_2 / ||b||_2 (rows are problems)."""
    num = np.linalg.norm(a - b, axis=1)
    den = np.linalg.norm(b, axis=1)
    den = np.where(den == 0.0, 1.0, den)
    return num / den


def assert_lqr_parity(gpu: dict, ref: dict, tol: float = REL_TOL, mask=None):
    for k in ("x", "u", "y"):
        if ref[k].shape[1] == 0:
            continue
        err = rel_err(gpu[k], ref[k])
        if mask is not None:
            err = err[mask]
        assert err.size == 0 or err.max() <= tol, (k, float(err.max()))


def gpu_lqr_factor_solve(s: pyoracle.Structure, host: dict, force_generic=False, fused=True,
                         pad_variable_dims=False, parallel_in_time=None):
    """Device path: pack -> (fused | factor + solve) -> unpack.  Returns dict + LQR."""
    dims, topo = to_structs(s)
    batch = host["q"].shape[0]
    lqr = LQR(dims, topo, batch, force_generic=force_generic,
              pad_variable_dims=pad_variable_dims, parallel_in_time=parallel_in_time)
    inp = lqr.pack_input(host)
    out = lqr.alloc_output()
    if fused:
        status = lqr.factor_solve(inp, out)
    else:
        status = lqr.factor_with_status(inp)
        lqr.solve(inp, out)
    res = lqr.unpack_output(out)
    res["status"] = status[:batch].cpu().numpy()
    norms, stats = lqr.residual(inp, out, status)
    res["residual"] = norms[:batch].cpu().numpy()
    res["stats"] = stats.cpu().numpy()
    return res, lqr
