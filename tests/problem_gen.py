"""Seeded synthetic problems for the parity tests (numpy, problem-major).

Distributions follow the reference's benchmark generators
(benchmarks/lqr_benchmark.cpp:61-96 and benchmarks/newton_kkt_benchmark.cpp:
171-240); the random stream itself is numpy's (the reference's
std::normal_distribution stream is implementation-defined).
"""
from __future__ import annotations

import numpy as np

from oracle.pyoracle import Structure, kkt_sizes


def lqr_benchmark_batch(n, m, T, batch, seed=0, dense_M=False):
    """Uniform chain; returns (Structure, dict of [batch, size] arrays)."""
    rng = np.random.default_rng(seed)
    s = Structure.chain(T, n, m)

    def spd(count, d, shift):
        Z = rng.standard_normal((batch, count, d, d))
        S = np.einsum("bckj,bcki->bcij", Z, Z)  # Z^T Z
        S = S + shift * np.eye(d)
        return S

    def colmajor(x):  # [batch, count, r, c] -> [batch, count * r * c], column-major blocks
        return np.ascontiguousarray(np.swapaxes(x, -1, -2)).reshape(batch, -1)

    A = 0.05 * rng.standard_normal((batch, T, n, n)) + np.eye(n)
    B = 0.1 * rng.standard_normal((batch, T, n, m))
    M = 0.01 * rng.standard_normal((batch, T, n, m)) if dense_M else np.zeros((batch, T, n, m))
    R = spd(T, m, 1.01)
    Q = spd(T + 1, n, 1e-3)
    d = dict(Q=colmajor(Q), M=colmajor(M), R=colmajor(R),
             q=rng.standard_normal((batch, (T + 1) * n)),
             r=rng.standard_normal((batch, T * m)), A=colmajor(A), B=colmajor(B),
             c=rng.standard_normal((batch, (T + 1) * n)),
             delta=1e-3 + 1e-1 * rng.random((batch, (T + 1) * n)))
    return s, d


def variable_tree_batch(s: Structure, batch, seed=0):
    """Any tree / per-node dims: well-conditioned random data."""
    rng = np.random.default_rng(seed)
    sd, cd = s.state_dims, s.control_dims
    N, E = len(sd), len(cd)

    def cm(x):
        return np.ascontiguousarray(np.swapaxes(x, -1, -2)).reshape(batch, -1)

    def spd(d, shift):
        Z = rng.standard_normal((batch, d, d))
        return np.einsum("bkj,bki->bij", Z, Z) + shift * np.eye(d)

    out = {k: [] for k in ("Q", "M", "R", "q", "r", "A", "B", "c", "delta")}
    for i in range(N):
        out["Q"].append(cm(spd(sd[i], 0.5)))
        out["q"].append(rng.standard_normal((batch, sd[i])))
        out["c"].append(rng.standard_normal((batch, sd[i])))
        out["delta"].append(0.05 + rng.random((batch, sd[i])))
    for e in range(E):
        npar, nch, m = sd[s.parents[e]], sd[s.children[e]], cd[e]
        out["M"].append(cm(0.05 * rng.standard_normal((batch, npar, m))))
        out["R"].append(cm(spd(m, 1.0)))
        out["r"].append(rng.standard_normal((batch, m)))
        out["A"].append(cm(0.3 * rng.standard_normal((batch, nch, npar))))
        out["B"].append(cm(0.3 * rng.standard_normal((batch, nch, m))))
    return {k: (np.concatenate(v, axis=1) if v else np.zeros((batch, 0))) for k, v in out.items()}


def newton_kkt_batch(s: Structure, batch, seed=0, r2_max=1e9):
    """newton_kkt_benchmark.cpp:171-240 distribution on any structure.

    Returns (model dict, w, r1, r2, r3, rhs), all problem-major.
    """
    rng = np.random.default_rng(seed)
    sd, cd = s.state_dims, s.control_dims
    N, E = len(sd), len(cd)
    z = lambda a, i: 0 if a is None else int(a[i])

    def cm(x):
        return np.ascontiguousarray(np.swapaxes(x, -1, -2)).reshape(batch, -1)

    def spd(d, shift):
        Z = rng.standard_normal((batch, d, d))
        return np.einsum("bkj,bki->bij", Z, Z) + shift * np.eye(d)

    m = {k: [] for k in ("node_hxx", "node_jc", "node_jg", "edge_hxx", "edge_hxu", "edge_huu",
                         "edge_A", "edge_B", "edge_jcx", "edge_jcu", "edge_jgx", "edge_jgu")}
    for i in range(N):
        n = sd[i]
        m["node_jc"].append(cm(0.1 * rng.standard_normal((batch, z(s.node_c, i), n))))
        m["node_jg"].append(cm(0.1 * rng.standard_normal((batch, z(s.node_g, i), n))))
        m["node_hxx"].append(cm(spd(n, 1e-3)))
    for e in range(E):
        npar, nch, mm = sd[s.parents[e]], sd[s.children[e]], cd[e]
        c, g = z(s.edge_c, e), z(s.edge_g, e)
        A = 0.05 * rng.standard_normal((batch, nch, npar))
        if nch == npar:
            A = A + np.eye(nch)
        m["edge_A"].append(cm(A))
        m["edge_B"].append(cm(0.1 * rng.standard_normal((batch, nch, mm))))
        m["edge_jcx"].append(cm(0.1 * rng.standard_normal((batch, c, npar))))
        m["edge_jcu"].append(cm(0.1 * rng.standard_normal((batch, c, mm))))
        m["edge_jgx"].append(cm(0.1 * rng.standard_normal((batch, g, npar))))
        m["edge_jgu"].append(cm(0.1 * rng.standard_normal((batch, g, mm))))
        m["edge_hxx"].append(np.zeros((batch, npar * npar)))
        m["edge_hxu"].append(cm(0.01 * rng.standard_normal((batch, npar, mm))))
        m["edge_huu"].append(cm(spd(mm, 1.0)))
    model = {k: (np.concatenate(v, axis=1) if v else np.zeros((batch, 0))) for k, v in m.items()}
    sz = kkt_sizes(s)

    def logu(shape, lo, hi):
        return np.exp(np.log(lo) + (np.log(hi) - np.log(lo)) * rng.random(shape))

    r1 = np.full((batch, sz["x_dim"]), 1e-8)
    r2 = logu((batch, sz["y_dim"]), 1e-3, r2_max)
    w = logu((batch, sz["z_dim"]), 1e-2, 1e3)
    r3 = logu((batch, sz["z_dim"]), 1e-3, 1e1)
    rhs = rng.standard_normal((batch, sz["kkt_dim"]))
    return model, w, r1, r2, r3, rhs


def newton_kkt_theta_batch(s: Structure, p: int, batch, seed=0, r2_max=1e9):
    """newton_kkt_batch plus the theta blocks of the reference's theta benchmark problems
    (newton_kkt_benchmark.cpp:180-225): every coupling block 1e-3 N(0,1), d2L_dtheta2 zero
    except Z'Z + 100 I on the last node.  Returns (model, theta, w, r1, r2, r3, rhs) with r1
    and rhs in the full layout [x_s, theta | y | z]."""
    from oracle.pyoracle import theta_sizes

    model, w, r1, r2, r3, rhs = newton_kkt_batch(s, batch, seed=seed, r2_max=r2_max)
    rng = np.random.default_rng(seed + 7919)
    sz = kkt_sizes(s)
    tz = theta_sizes(s, p)
    theta = {k: 1e-3 * rng.standard_normal((batch, n)) for k, n in tz.items()}
    N = len(s.state_dims)
    Z = rng.standard_normal((batch, p, p))
    last = np.einsum("bkj,bki->bij", Z, Z) + 100.0 * np.eye(p)
    htt = np.zeros((batch, N, p * p))
    htt[:, N - 1] = np.ascontiguousarray(np.swapaxes(last, -1, -2)).reshape(batch, -1)
    theta["node_htt"] = htt.reshape(batch, -1)
    theta["edge_htt"] = np.zeros((batch, tz["edge_htt"]))
    sx = sz["x_dim"]
    r1_full = np.concatenate([r1, np.full((batch, p), 1e-8)], axis=1)
    rhs_full = np.concatenate([rhs[:, :sx], rng.standard_normal((batch, p)), rhs[:, sx:]], axis=1)
    return model, theta, w, r1_full, r2, r3, rhs_full
