"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end over oracle/liboracle.so (the Eigen-free CPU restatement of the
reference's lqr.cpp / helpers.cpp, see riccati_oracle.hpp).  Imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; the product package never imports it.

All batched arrays are problem-major: ``X[batch, flat_per_problem_index]``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)
_c_i64_p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def _ip(a):
    return None if a is None else a.ctypes.data_as(_c_int_p)


def _dp(a):
    return None if a is None else a.ctypes.data_as(_c_dbl_p)


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _f64(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a


@dataclass
class Structure:
    """Topology + dimensions shared by every problem of a batch."""

    parents: np.ndarray
    children: np.ndarray
    root: int
    state_dims: np.ndarray
    control_dims: np.ndarray
    node_c: np.ndarray | None = None
    node_g: np.ndarray | None = None
    edge_c: np.ndarray | None = None
    edge_g: np.ndarray | None = None

    def __post_init__(self):
        self.parents = _i32(self.parents)
        self.children = _i32(self.children)
        self.state_dims = _i32(self.state_dims)
        self.control_dims = _i32(self.control_dims)
        for name in ("node_c", "node_g", "edge_c", "edge_g"):
            v = getattr(self, name)
            if v is not None:
                setattr(self, name, _i32(v))

    @property
    def num_edges(self) -> int:
        return int(self.parents.shape[0])

    @staticmethod
    def chain(num_edges, n, m, **kw):
        e = np.arange(num_edges, dtype=np.int32)
        sd = np.full(num_edges + 1, n, np.int32) if np.isscalar(n) else _i32(n)
        cd = np.full(num_edges, m, np.int32) if np.isscalar(m) else _i32(m)
        return Structure(e, e + 1, 0, sd, cd, **kw)

    def _topo_args(self):
        return (ctypes.c_int(self.num_edges), ctypes.c_int(self.root), _ip(self.parents),
                _ip(self.children))

    def _dim_args(self):
        return (_ip(self.state_dims), _ip(self.control_dims))

    def _cg_args(self):
        return (_ip(self.node_c), _ip(self.node_g), _ip(self.edge_c), _ip(self.edge_g))


def compile_topology(s: Structure):
    """Returns (status, child_offsets, child_edges, preorder, postorder)."""
    E = s.num_edges
    co = np.zeros(E + 2, np.int32)
    ce = np.zeros(max(E, 1), np.int32)
    pre = np.zeros(E + 1, np.int32)
    post = np.zeros(E + 1, np.int32)
    st = lib().oracle_compile_topology(*s._topo_args(), _ip(co), _ip(ce), _ip(pre), _ip(post))
    return int(st), co, ce[:E], pre, post


LQR_INPUT_NAMES = ("Q", "M", "R", "q", "r", "A", "B", "c", "delta")


def lqr_sizes(s: Structure) -> dict:
    """Per-problem element counts of every flat LQR array."""
    out = np.zeros(7, np.int64)
    lib().oracle_lqr_sizes(*s._topo_args(), *s._dim_args(), out.ctypes.data_as(_c_i64_p))
    nn, n, nm, mm, m, a, b = (int(v) for v in out)
    return dict(Q=nn, M=nm, R=mm, q=n, r=m, A=a, B=b, c=n, delta=n, x=n, u=m, y=n)


def lqr_factor_solve(s: Structure, inputs: dict, solve: bool = True, residual: bool = True,
                     repeats: int = 1, nthreads: int = 0):
    """Batched factor(+solve).  inputs[name] is [batch, size(name)].

    Returns dict(x, u, y, status, residual, seconds).
    """
    sz = lqr_sizes(s)
    arrs = {k: _f64(inputs[k]) for k in LQR_INPUT_NAMES}
    batch = arrs["q"].shape[0]
    for k in LQR_INPUT_NAMES:
        assert arrs[k].shape == (batch, sz[k]), (k, arrs[k].shape, (batch, sz[k]))
    x = np.zeros((batch, sz["x"]))
    u = np.zeros((batch, sz["u"]))
    y = np.zeros((batch, sz["y"]))
    status = np.zeros(batch, np.int32)
    res = np.zeros(batch) if residual else None
    secs = ctypes.c_double(0.0)
    lib().oracle_lqr_batch(
        *s._topo_args(), *s._dim_args(), ctypes.c_int64(batch),
        *[_dp(arrs[k]) for k in LQR_INPUT_NAMES], _dp(x), _dp(u), _dp(y), _ip(status),
        _dp(res), ctypes.c_int(3 if solve else 1), ctypes.c_int(repeats),
        ctypes.c_int(nthreads), ctypes.byref(secs))
    return dict(x=x, u=u, y=y, status=status, residual=res, seconds=secs.value)


def lqr_residual(s: Structure, inputs: dict, x, u, y) -> np.ndarray:
    arrs = {k: _f64(inputs[k]) for k in LQR_INPUT_NAMES}
    x, u, y = _f64(x), _f64(u), _f64(y)
    batch = x.shape[0]
    res = np.zeros(batch)
    st = lib().oracle_lqr_residual_batch(
        *s._topo_args(), *s._dim_args(), ctypes.c_int64(batch),
        *[_dp(arrs[k]) for k in LQR_INPUT_NAMES], _dp(x), _dp(u), _dp(y), _dp(res))
    assert st == 0
    return res


KKT_MODEL_NAMES = ("node_hxx", "node_jc", "node_jg", "edge_hxx", "edge_hxu", "edge_huu",
                   "edge_A", "edge_B", "edge_jcx", "edge_jcu", "edge_jgx", "edge_jgu")


def kkt_sizes(s: Structure) -> dict:
    out = np.zeros(16, np.int64)
    lib().oracle_kkt_sizes(*s._topo_args(), *s._dim_args(), *s._cg_args(),
                           out.ctypes.data_as(_c_i64_p))
    d = dict(x_dim=int(out[0]), y_dim=int(out[1]), z_dim=int(out[2]))
    for i, k in enumerate(KKT_MODEL_NAMES):
        d[k] = int(out[3 + i])
    d["kkt_dim"] = d["x_dim"] + d["y_dim"] + d["z_dim"]
    return d


def kkt_offsets(s: Structure) -> dict:
    E = s.num_edges
    names_n = ("x_state", "y_dyn", "y_node_c", "z_node")
    names_e = ("x_control", "y_edge_c", "z_edge")
    o = {k: np.zeros(E + 1, np.int32) for k in names_n}
    o.update({k: np.zeros(max(E, 1), np.int32) for k in names_e})
    lib().oracle_kkt_offsets(*s._topo_args(), *s._dim_args(), *s._cg_args(),
                             _ip(o["x_state"]), _ip(o["x_control"]), _ip(o["y_dyn"]),
                             _ip(o["y_node_c"]), _ip(o["y_edge_c"]), _ip(o["z_node"]),
                             _ip(o["z_edge"]))
    for k in names_e:
        o[k] = o[k][:E]
    return o


def kkt_factor_solve(s: Structure, model: dict, w, r1, r2, r3, b, solve=True, residual=True,
                     repeats: int = 1, nthreads: int = 0):
    """Batched CallbackProvider::factor (+ solve, + ||K sol - b||)."""
    sz = kkt_sizes(s)
    mdl = {k: _f64(model[k]) for k in KKT_MODEL_NAMES}
    w, r1, r2, r3, b = (_f64(v) for v in (w, r1, r2, r3, b))
    batch = b.shape[0]
    for k in KKT_MODEL_NAMES:
        assert mdl[k].shape == (batch, sz[k]), (k, mdl[k].shape, (batch, sz[k]))
    assert w.shape == (batch, sz["z_dim"]) and r3.shape == w.shape
    assert r1.shape == (batch, sz["x_dim"]) and r2.shape == (batch, sz["y_dim"])
    assert b.shape == (batch, sz["kkt_dim"])
    sol = np.zeros_like(b)
    ok = np.zeros(batch, np.int32)
    lqr_status = np.zeros(batch, np.int32)
    res = np.zeros(batch) if residual else None
    secs = ctypes.c_double(0.0)
    mode = 1 | (2 if solve else 0) | (4 if (residual and solve) else 0)
    lib().oracle_kkt_batch(
        *s._topo_args(), *s._dim_args(), *s._cg_args(), ctypes.c_int64(batch),
        *[_dp(mdl[k]) for k in KKT_MODEL_NAMES], _dp(w), _dp(r1), _dp(r2), _dp(r3), _dp(b),
        _dp(sol), _ip(ok), _ip(lqr_status), _dp(res), ctypes.c_int(mode),
        ctypes.c_int(repeats), ctypes.c_int(nthreads), ctypes.byref(secs))
    return dict(sol=sol, ok=ok, lqr_status=lqr_status, residual=res, seconds=secs.value)


def kkt_apply(s: Structure, model: dict, w, r1, r2, r3, x, y=None) -> np.ndarray:
    """y += K(w, r1, r2, r3) x  (add_Kx_to_y); x, y are [batch, kkt_dim]."""
    mdl = {k: _f64(model[k]) for k in KKT_MODEL_NAMES}
    w, r1, r2, r3, x = (_f64(v) for v in (w, r1, r2, r3, x))
    y = np.zeros_like(x) if y is None else _f64(y).copy()
    st = lib().oracle_kkt_apply_batch(
        *s._topo_args(), *s._dim_args(), *s._cg_args(), ctypes.c_int64(x.shape[0]),
        *[_dp(mdl[k]) for k in KKT_MODEL_NAMES], _dp(w), _dp(r1), _dp(r2), _dp(r3), _dp(x),
        _dp(y))
    assert st == 0
    return y


THETA_MODEL_NAMES = ("node_hxt", "node_jct", "node_jgt", "node_htt", "edge_hxt", "edge_hut",
                     "edge_dynt", "edge_jct", "edge_jgt", "edge_htt")


def theta_sizes(s: Structure, p: int) -> dict:
    """Per-problem element counts of the ten theta block arrays (theta_oracle.hpp)."""
    out = np.zeros(10, np.int64)
    lib().oracle_kkt_theta_sizes(*s._topo_args(), *s._dim_args(), *s._cg_args(),
                                 ctypes.c_int(p), out.ctypes.data_as(_c_i64_p))
    return {k: int(out[i]) for i, k in enumerate(THETA_MODEL_NAMES)}


def _theta_call(s, p, model, theta, w, r1, r2, r3, vec_in, vec_out, ok, mode):
    mdl = [_f64(model[k]) for k in KKT_MODEL_NAMES]
    th = [_f64(theta[k]) for k in THETA_MODEL_NAMES]
    tz = theta_sizes(s, p)
    batch = vec_in.shape[0]
    for k, a in zip(THETA_MODEL_NAMES, th):
        assert a.shape == (batch, tz[k]), (k, a.shape, (batch, tz[k]))
    arr_m = (_c_dbl_p * 12)(*[_dp(a) for a in mdl])
    arr_t = (_c_dbl_p * 10)(*[_dp(a) for a in th])
    st = lib().oracle_kkt_theta_batch(
        *s._topo_args(), *s._dim_args(), *s._cg_args(), ctypes.c_int(p), ctypes.c_int64(batch),
        arr_m, arr_t, _dp(w), _dp(r1), _dp(r2), _dp(r3), _dp(vec_in), _dp(vec_out), _ip(ok),
        ctypes.c_int(mode))
    assert st == 0
    return mdl, th  # keep the arrays alive until the call has returned


def kkt_theta_factor_solve(s: Structure, p: int, model: dict, theta: dict, w, r1, r2, r3, b,
                           solve=True):
    """CallbackProvider::factor (+ solve) with theta_dim = p; vectors in the full layout
    [x_s, theta | y | z], r1 of length x_dim + p."""
    sz = kkt_sizes(s)
    w, r1, r2, r3, b = (_f64(v) for v in (w, r1, r2, r3, b))
    batch = b.shape[0]
    assert r1.shape == (batch, sz["x_dim"] + p) and b.shape == (batch, sz["kkt_dim"] + p)
    sol = np.zeros_like(b)
    ok = np.zeros(batch, np.int32)
    _theta_call(s, p, model, theta, w, r1, r2, r3, b, sol, ok, 1 | (2 if solve else 0))
    return dict(sol=sol, ok=ok)


def kkt_theta_apply(s: Structure, p: int, model: dict, theta: dict, w, r1, r2, r3, x, y=None):
    """y += K x with the theta rows / columns (add_Kx_to_y, theta_dim = p)."""
    w, r1, r2, r3, x = (_f64(v) for v in (w, r1, r2, r3, x))
    y = np.zeros_like(x) if y is None else _f64(y).copy()
    ok = np.zeros(x.shape[0], np.int32)
    _theta_call(s, p, model, theta, w, r1, r2, r3, x, y, ok, 8)
    return y


MODEL_VALUE_NAMES = ("node_f", "node_df_dx", "node_df_dtheta", "node_c", "node_g", "edge_f",
                     "edge_df_dx", "edge_df_du", "edge_df_dtheta", "edge_dyn_res", "edge_c",
                     "edge_g")


def model_value_sizes(s: Structure, p: int = 0) -> dict:
    """Per-problem element counts of the twelve value arrays of one model evaluation."""
    E, N = s.num_edges, s.num_edges + 1
    n = np.asarray(s.state_dims)
    par, chi = np.asarray(s.parents)[:E], np.asarray(s.children)[:E]
    return dict(node_f=N, node_df_dx=int(n.sum()), node_df_dtheta=N * p,
                node_c=int(np.sum(s.node_c)), node_g=int(np.sum(s.node_g)), edge_f=E,
                edge_df_dx=int(n[par].sum()) if E else 0,
                edge_df_du=int(np.sum(np.asarray(s.control_dims)[:E])), edge_df_dtheta=E * p,
                edge_dyn_res=int(n[chi].sum()) if E else 0,
                edge_c=int(np.sum(np.asarray(s.edge_c)[:E])),
                edge_g=int(np.sum(np.asarray(s.edge_g)[:E])))


def model_scatter(s: Structure, values: dict, x, initial_state, p: int = 0, new_x: bool = True):
    """The model callback's scatter (sip_optimal_control.cpp:44-123): f, gradient_f, c, g."""
    sz, ksz = model_value_sizes(s, p), kkt_sizes(s)
    vals = [_f64(values[k]) for k in MODEL_VALUE_NAMES]
    x, x0 = _f64(x), _f64(initial_state)
    batch = x.shape[0]
    for k, a in zip(MODEL_VALUE_NAMES, vals):
        assert a.shape == (batch, sz[k]), (k, a.shape, (batch, sz[k]))
    assert x.shape == (batch, ksz["x_dim"] + p)
    f = np.zeros(batch)
    grad = np.zeros((batch, ksz["x_dim"] + p))
    c = np.zeros((batch, ksz["y_dim"]))
    g = np.zeros((batch, ksz["z_dim"]))
    arr = (_c_dbl_p * 12)(*[_dp(a) for a in vals])
    st = lib().oracle_model_scatter_batch(
        *s._topo_args(), *s._dim_args(), *s._cg_args(), ctypes.c_int(p), ctypes.c_int64(batch),
        arr, _dp(x), _dp(x0), ctypes.c_int(1 if new_x else 0), _dp(f), _dp(grad), _dp(c), _dp(g))
    assert st == 0
    return dict(f=f, gradient_f=grad, c=c, g=g)


def max_threads() -> int:
    return int(lib().oracle_max_threads())
