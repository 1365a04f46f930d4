"""Closed-form fixtures of the reference's own tests, restated as flat arrays.

Every builder cites the reference lines that define the numbers.  Eigen's
comma initialiser (``M << a, b, c, ...``) fills ROW by row, so those matrices
are written here row-major and flattened column-major (``order="F"``), which
is the memory order the reference hands to the solver.
"""
from __future__ import annotations

import numpy as np

from oracle.pyoracle import Structure


def pack(blocks):
    """Concatenate matrices / vectors column-major into one flat array."""
    if len(blocks) == 0:
        return np.zeros(0)
    return np.concatenate([np.asarray(b, dtype=np.float64).flatten(order="F") for b in blocks])


def pack_problem(Q, M, R, q, r, A, B, c, delta):
    """One problem -> dict of [1, size] arrays (problem-major batch of one)."""
    d = dict(Q=pack(Q), M=pack(M), R=pack(R), q=pack(q), r=pack(r), A=pack(A), B=pack(B),
             c=pack(c), delta=pack(delta))
    return {k: v[None, :] for k, v in d.items()}


def linspaced(n, lo, hi):
    # Eigen::VectorXd::LinSpaced(n, lo, hi); size 1 yields `hi` (Eigen 3.4 docs).
    if n == 1:
        return np.array([hi])
    return np.linspace(lo, hi, n)


# --- tests/lqr_test.cpp:45-75 : identity problem used by the status tests -----------------
def identity_chain(n, m, T):
    s = Structure.chain(T, n, m)
    Q = [np.eye(n) for _ in range(T + 1)]
    q = [np.zeros(n) for _ in range(T + 1)]
    c = [np.zeros(n) for _ in range(T + 1)]
    delta = [np.ones(n) for _ in range(T + 1)]
    M = [np.zeros((n, m)) for _ in range(T)]
    R = [np.eye(m) for _ in range(T)]
    A = [np.eye(n) for _ in range(T)]
    B = [np.ones((n, m)) for _ in range(T)]
    r = [np.zeros(m) for _ in range(T)]
    return s, dict(Q=Q, M=M, R=R, q=q, r=r, A=A, B=B, c=c, delta=delta)


# --- tests/lqr_test.cpp:229-247 : chain n=3, m=2, T=3, non-uniform diagonal delta ----------
def nonuniform_delta_chain():
    n, m, T = 3, 2, 3
    s, p = identity_chain(n, m, T)
    for i in range(T):
        p["A"][i] = np.array([[1.0 + 0.02 * i, 0.03, -0.01],
                              [-0.02, 0.95 + 0.01 * i, 0.04],
                              [0.01, -0.03, 1.02 - 0.01 * i]])
        p["B"][i] = np.array([[0.2, -0.1], [0.05, 0.15], [-0.1, 0.08]])
        p["Q"][i] = np.diag([1.0 + 0.1 * i, 1.4 + 0.05 * i, 1.8 + 0.03 * i])
        p["R"][i] = np.diag([1.2 + 0.1 * i, 1.6 + 0.07 * i])
        p["q"][i] = np.array([0.2 + 0.01 * i, -0.1 + 0.02 * i, 0.05 - 0.03 * i])
        p["r"][i] = np.array([-0.2 + 0.03 * i, 0.1 - 0.01 * i])
        p["c"][i] = np.array([0.03 + 0.01 * i, -0.04 + 0.02 * i, 0.02 - 0.01 * i])
        p["delta"][i] = np.array([0.03 + 0.01 * i, 0.11 + 0.02 * i, 0.19 + 0.03 * i])
    p["Q"][T] = np.diag([1.3, 1.7, 2.1])
    p["q"][T] = np.array([0.06, -0.08, 0.12])
    p["c"][T] = np.array([-0.02, 0.05, -0.01])
    p["delta"][T] = np.array([0.07, 0.17, 0.29])
    return s, p


# --- tests/lqr_test.cpp:265-335 : 3-node branching tree, uniform dims ----------------------
def branch_tree():
    s = Structure([0, 0], [1, 2], 0, [2, 2, 2], [1, 1])
    Q = [np.array([[2.0, 0.1], [0.1, 1.5]]), np.array([[1.3, 0.2], [0.2, 1.7]]),
         np.array([[1.8, -0.1], [-0.1, 1.4]])]
    M = [np.array([[0.2], [-0.1]]), np.array([[-0.15], [0.05]])]
    R = [np.array([[1.6]]), np.array([[1.9]])]
    A = [np.array([[1.0, 0.2], [0.0, 0.9]]), np.array([[0.8, -0.1], [0.3, 1.1]])]
    B = [np.array([[0.4], [0.2]]), np.array([[-0.1], [0.5]])]
    q = [np.array([0.3, -0.2]), np.array([-0.1, 0.4]), np.array([0.2, 0.1])]
    r = [np.array([-0.3]), np.array([0.25])]
    c = [np.array([0.1, -0.2]), np.array([-0.05, 0.1]), np.array([0.2, 0.15])]
    delta = [np.array([0.7, 0.9]), np.array([0.8, 1.1]), np.array([1.0, 0.6])]
    return s, dict(Q=Q, M=M, R=R, q=q, r=r, A=A, B=B, c=c, delta=delta)


# --- tests/lqr_test.cpp:466-532 : variable-dimension branching tree ------------------------
def variable_dim_branch_tree():
    s = Structure([0, 0], [1, 2], 0, [2, 1, 3], [2, 1])
    Q = [np.array([[2.0, 0.1], [0.1, 1.7]]), np.array([[1.3]]),
         np.array([[1.8, 0.1, -0.2], [0.1, 1.6, 0.05], [-0.2, 0.05, 2.1]])]
    M = [np.array([[0.1, -0.2], [0.05, 0.15]]), np.array([[-0.1], [0.2]])]
    R = [np.array([[1.8, 0.1], [0.1, 1.5]]), np.array([[1.4]])]
    A = [np.array([[0.8, -0.3]]), np.array([[1.0, 0.2], [-0.1, 0.7], [0.3, -0.4]])]
    B = [np.array([[0.4, -0.2]]), np.array([[0.2], [-0.1], [0.5]])]
    q = [np.array([0.2, -0.15]), np.array([-0.05]), np.array([0.1, -0.2, 0.05])]
    r = [np.array([-0.1, 0.25]), np.array([-0.2])]
    c = [np.array([0.05, -0.1]), np.array([0.12]), np.array([-0.02, 0.04, -0.08])]
    delta = [np.array([0.8, 1.1]), np.array([0.9]), np.array([0.7, 1.0, 1.2])]
    return s, dict(Q=Q, M=M, R=R, q=q, r=r, A=A, B=B, c=c, delta=delta)


# --- tests/lqr_test.cpp:661-762 : five-node variable-dimension tree ------------------------
def five_node_tree(parents=(0, 0, 1, 1), children=(1, 2, 3, 4)):
    state_dims = [3, 1, 2, 4, 2]
    control_dims = [2, 1, 3, 1]
    s = Structure(list(parents), list(children), 0, state_dims, control_dims)
    Q, q, c, delta = [], [], [], []
    for node in range(5):
        n = state_dims[node]
        Qn = np.eye(n) * (1.5 + 0.2 * node)
        for col in range(n):
            for row in range(col + 1, n):
                Qn[row, col] = 0.02 * (row + col + node + 1)
                Qn[col, row] = Qn[row, col]
        Q.append(Qn)
        q.append(linspaced(n, -0.15 + 0.03 * node, 0.12 + 0.02 * node))
        c.append(linspaced(n, 0.05 * node, 0.04 + 0.03 * node))
        delta.append(linspaced(n, 0.7 + 0.05 * node, 1.0 + 0.04 * node))
    M, A, B, R, r = [], [], [], [], []
    # The data is defined with the fixture's original (valid) topology.
    base_parent, base_child = (0, 0, 1, 1), (1, 2, 3, 4)
    for edge in range(4):
        n_parent = state_dims[base_parent[edge]]
        n_child = state_dims[base_child[edge]]
        m = control_dims[edge]
        Me = np.zeros((n_parent, m))
        Ae = np.zeros((n_child, n_parent))
        Be = np.zeros((n_child, m))
        for col in range(m):
            for row in range(n_parent):
                Me[row, col] = 0.015 * ((edge + 1) * (row + 1) - col)
        for col in range(n_parent):
            for row in range(n_child):
                Ae[row, col] = 0.08 * (row + 1) / (edge + col + 2)
        for col in range(m):
            for row in range(n_child):
                Be[row, col] = -0.06 * (col + 1) / (edge + row + 2)
        Re = np.eye(m) * (1.8 + 0.1 * edge)
        for col in range(m):
            for row in range(col + 1, m):
                Re[row, col] = 0.03 * (row + col + 1)
                Re[col, row] = Re[row, col]
        M.append(Me)
        A.append(Ae)
        B.append(Be)
        R.append(Re)
        r.append(linspaced(m, -0.2 + 0.04 * edge, 0.1 + 0.03 * edge))
    return s, dict(Q=Q, M=M, R=R, q=q, r=r, A=A, B=B, c=c, delta=delta)


def dense_kkt_solve(s: Structure, p: dict):
    """Assemble the KKT system the LQR solves and solve it densely.

    Same equations as tests/lqr_test.cpp:859-929 (solve_dense_kkt), unknowns
    ordered [x nodes | u edges | y nodes]; numpy's LU replaces
    colPivHouseholderQr.  Returns (x list, u list, y list).
    """
    sd, cd = s.state_dims, s.control_dims
    N, E = len(sd), len(cd)
    xo = np.concatenate([[0], np.cumsum(sd)])
    uo = xo[-1] + np.concatenate([[0], np.cumsum(cd)])
    yo = uo[-1] + np.concatenate([[0], np.cumsum(sd)])
    dim = int(yo[-1])
    Kmat = np.zeros((dim, dim))
    rhs = np.zeros(dim)
    row = 0
    for node in range(N):
        n = sd[node]
        Kmat[row:row + n, xo[node]:xo[node] + n] += p["Q"][node]
        Kmat[row:row + n, yo[node]:yo[node] + n] -= np.eye(n)
        for e in range(E):
            if s.parents[e] == node:
                ch = s.children[e]
                Kmat[row:row + n, uo[e]:uo[e] + cd[e]] += p["M"][e]
                Kmat[row:row + n, yo[ch]:yo[ch] + sd[ch]] += p["A"][e].T
        rhs[row:row + n] = -p["q"][node]
        row += n
    for e in range(E):
        pa, ch, m = s.parents[e], s.children[e], cd[e]
        Kmat[row:row + m, xo[pa]:xo[pa] + sd[pa]] += p["M"][e].T
        Kmat[row:row + m, uo[e]:uo[e] + m] += p["R"][e]
        Kmat[row:row + m, yo[ch]:yo[ch] + sd[ch]] += p["B"][e].T
        rhs[row:row + m] = -p["r"][e]
        row += m
    root = s.root
    n = sd[root]
    Kmat[row:row + n, xo[root]:xo[root] + n] -= np.eye(n)
    Kmat[row:row + n, yo[root]:yo[root] + n] -= np.diag(p["delta"][root])
    rhs[row:row + n] = -p["c"][root]
    row += n
    for e in range(E):
        pa, ch = s.parents[e], s.children[e]
        nc = sd[ch]
        Kmat[row:row + nc, xo[pa]:xo[pa] + sd[pa]] += p["A"][e]
        Kmat[row:row + nc, uo[e]:uo[e] + cd[e]] += p["B"][e]
        Kmat[row:row + nc, xo[ch]:xo[ch] + nc] -= np.eye(nc)
        Kmat[row:row + nc, yo[ch]:yo[ch] + nc] -= np.diag(p["delta"][ch])
        rhs[row:row + nc] = -p["c"][ch]
        row += nc
    assert row == dim
    sol = np.linalg.solve(Kmat, rhs)
    x = [sol[xo[i]:xo[i + 1]] for i in range(N)]
    u = [sol[uo[e]:uo[e + 1]] for e in range(E)]
    y = [sol[yo[i]:yo[i + 1]] for i in range(N)]
    return x, u, y


# --- tests/variable_dimensions_test.cpp:46-133 : Newton-KKT model fixture ------------------
def fill_sequence(size, scale):
    return scale * np.arange(1, size + 1, dtype=np.float64)  # :46-50


def kkt_model(s: Structure):
    """initialize_model (:77-133) with theta_dim == 0; flat [1, size] arrays."""
    sd, cd = s.state_dims, s.control_dims
    N, E = len(sd), len(cd)
    nc = s.node_c if s.node_c is not None else np.zeros(N, np.int32)
    ng = s.node_g if s.node_g is not None else np.zeros(N, np.int32)
    ec = s.edge_c if s.edge_c is not None else np.zeros(E, np.int32)
    eg = s.edge_g if s.edge_g is not None else np.zeros(E, np.int32)
    m = {k: [] for k in ("node_hxx", "node_jc", "node_jg", "edge_hxx", "edge_hxu", "edge_huu",
                         "edge_A", "edge_B", "edge_jcx", "edge_jcu", "edge_jgx", "edge_jgu")}
    for node in range(N):
        n = sd[node]
        m["node_jc"].append(fill_sequence(nc[node] * n, 0.013 * (node + 1)))
        m["node_jg"].append(fill_sequence(ng[node] * n, -0.011 * (node + 1)))
        m["node_hxx"].append((np.eye(n) * (2.5 + 0.2 * node)).flatten(order="F"))
    for e in range(E):
        npar, nch, mm = sd[s.parents[e]], sd[s.children[e]], cd[e]
        m["edge_A"].append(fill_sequence(nch * npar, 0.025 + 0.004 * e))
        m["edge_B"].append(fill_sequence(nch * mm, -0.031 - 0.003 * e))
        m["edge_jcx"].append(fill_sequence(ec[e] * npar, 0.017 * (e + 1)))
        m["edge_jcu"].append(fill_sequence(ec[e] * mm, 0.019 * (e + 1)))
        m["edge_jgx"].append(fill_sequence(eg[e] * npar, -0.014 * (e + 1)))
        m["edge_jgu"].append(fill_sequence(eg[e] * mm, 0.016 * (e + 1)))
        m["edge_hxx"].append((np.eye(npar) * (0.3 + 0.05 * e)).flatten(order="F"))
        m["edge_hxu"].append(fill_sequence(npar * mm, 0.009 * (e + 1)))
        m["edge_huu"].append((np.eye(mm) * (3.0 + 0.2 * e)).flatten(order="F"))
    return {k: (np.concatenate(v) if len(v) else np.zeros(0))[None, :] for k, v in m.items()}


def kkt_theta_model(s: Structure, p: int, theta_diagonal: float):
    """The theta blocks of initialize_model (:77-133): flat [1, size] arrays in
    oracle.pyoracle.THETA_MODEL_NAMES order."""
    sd, cd = s.state_dims, s.control_dims
    N, E = len(sd), len(cd)
    nc = s.node_c if s.node_c is not None else np.zeros(N, np.int32)
    ng = s.node_g if s.node_g is not None else np.zeros(N, np.int32)
    ec = s.edge_c if s.edge_c is not None else np.zeros(E, np.int32)
    eg = s.edge_g if s.edge_g is not None else np.zeros(E, np.int32)
    t = {k: [] for k in ("node_hxt", "node_jct", "node_jgt", "node_htt", "edge_hxt", "edge_hut",
                         "edge_dynt", "edge_jct", "edge_jgt", "edge_htt")}
    htt = (np.eye(p) * theta_diagonal).flatten(order="F")
    for node in range(N):
        t["node_jct"].append(fill_sequence(nc[node] * p, 0.001 * (node + 1)))
        t["node_jgt"].append(fill_sequence(ng[node] * p, -0.0007 * (node + 1)))
        t["node_hxt"].append(fill_sequence(sd[node] * p, 0.0005 * (node + 1)))
        t["node_htt"].append(htt)
    for e in range(E):
        npar, nch, mm = sd[s.parents[e]], sd[s.children[e]], cd[e]
        t["edge_dynt"].append(fill_sequence(nch * p, 0.0009 * (e + 1)))
        t["edge_jct"].append(fill_sequence(ec[e] * p, 0.0008 * (e + 1)))
        t["edge_jgt"].append(fill_sequence(eg[e] * p, -0.0006 * (e + 1)))
        t["edge_hxt"].append(fill_sequence(npar * p, 0.0004 * (e + 1)))
        t["edge_hut"].append(fill_sequence(mm * p, -0.0003 * (e + 1)))
        t["edge_htt"].append(htt)
    return {k: (np.concatenate(v) if len(v) else np.zeros(0))[None, :] for k, v in t.items()}


def kkt_case_schur():  # :338-363, theta_dim = 2, theta_diagonal = 6.0, tolerance 1e-8
    s = Structure([0, 0], [1, 2], 0, [2, 1, 3], [1, 2], node_c=[1, 0, 1], node_g=[0, 1, 1],
                  edge_c=[1, 2], edge_g=[2, 1])
    return s, 2, 6.0


def kkt_regularization(x_dim, y_dim, z_dim):
    """expect_kkt_solve (:143-155): w=1.3, r2=0.9, r3=0.4, r1=0.2+0.03(i+1), rhs=0.01(i+1)."""
    w = np.full((1, z_dim), 1.3)
    r2 = np.full((1, y_dim), 0.9)
    r3 = np.full((1, z_dim), 0.4)
    r1 = (fill_sequence(x_dim, 0.03) + 0.2)[None, :]
    rhs = fill_sequence(x_dim + y_dim + z_dim, 0.01)[None, :]
    return w, r1, r2, r3, rhs


# Cases of tests/variable_dimensions_test.cpp:265-336 (theta_dim == 0).
def kkt_case_chain():  # :265-290
    return Structure([0, 1], [1, 2], 0, [2, 1, 3], [1, 2], node_c=[1, 0, 2], node_g=[0, 2, 1],
                     edge_c=[1, 2], edge_g=[2, 1])


def kkt_case_siblings():  # :292-314
    return Structure([0, 0], [1, 2], 0, [2, 1, 3], [1, 2], node_c=[1, 0, 1], node_g=[1, 1, 0],
                     edge_c=[2, 1], edge_g=[1, 2])


def kkt_case_zero_dim_root():  # :316-336
    return Structure([0, 0], [1, 2], 0, [0, 1, 3], [1, 2], node_c=[0, 0, 0], node_g=[0, 0, 0],
                     edge_c=[0, 0], edge_g=[0, 0])
