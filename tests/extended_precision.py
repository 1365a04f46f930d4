"""Extended-precision truth for the regularized-LQR recursion on uniform chains.

numpy's longdouble (x87 80-bit, 64-bit mantissa, eps 1.1e-19) restatement of the reference's
statement sequence (lqr.cpp:645-731 factor, :738-870 solve).  It answers the question the FP64
parity tests cannot: two FP64 implementations that agree with each other to 1e-13 could both
be 1e-6 away from the exact answer at ill-conditioned regularization (delta up to 1e9, the
Newton-KKT benchmark's r2 range).  The same algorithm carried out with 2 000x smaller rounding
errors is a truth good to ~1e-3 of any FP64 implementation's own error, so `|fp64 - truth|`
IS that implementation's error.  Test infrastructure only (pure Python loops, tiny sizes)."""
import numpy as np

LD = np.longdouble


def _chol(A):
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        d = A[j, j] - L[j, :j] @ L[j, :j]
        assert d > 0
        L[j, j] = np.sqrt(d)
        for i in range(j + 1, n):
            L[i, j] = (A[i, j] - L[i, :j] @ L[j, :j]) / L[j, j]
    return L


def _solve_lower(L, B):
    X = np.array(B, dtype=LD, copy=True)
    for i in range(L.shape[0]):
        X[i] = (X[i] - L[i, :i] @ X[:i]) / L[i, i]
    return X


def _solve_upper(U, B):
    X = np.array(B, dtype=LD, copy=True)
    for i in reversed(range(U.shape[0])):
        X[i] = (X[i] - U[i, i + 1:] @ X[i + 1:]) / U[i, i]
    return X


def _chol_solve(L, B):
    return _solve_upper(L.T, _solve_lower(L, B))


def lqr_chain_truth(n, m, T, flat):
    """One problem: flat[name] are the per-problem flat arrays of tests/problem_gen.py
    (column-major blocks).  Returns x [(T+1) n], u [T m], y [(T+1) n] as float64."""
    g = {k: np.asarray(v, dtype=LD) for k, v in flat.items()}
    mat = lambda a, k, r, c: a[k * r * c:(k + 1) * r * c].reshape(c, r).T
    Q = [mat(g["Q"], k, n, n) for k in range(T + 1)]
    M = [mat(g["M"], k, n, m) for k in range(T)]
    R = [mat(g["R"], k, m, m) for k in range(T)]
    A = [mat(g["A"], k, n, n) for k in range(T)]
    B = [mat(g["B"], k, n, m) for k in range(T)]
    vec = lambda a, k, r: a[k * r:(k + 1) * r]
    I = np.eye(n, dtype=LD)
    V, W, LF, sd = [None] * (T + 1), [None] * (T + 1), [None] * (T + 1), [None] * (T + 1)
    K, LG = [None] * T, [None] * T

    def node(k, Vk):  # lqr.cpp:475-529, 722-727
        d = vec(g["delta"], k, n)
        assert (d > 0).all()
        sd[k] = np.sqrt(d)
        F = I + sd[k][:, None] * Vk * sd[k][None, :]
        LF[k] = _chol(F)
        Finv = _chol_solve(LF[k], I)
        W[k] = (I - Finv) / sd[k][:, None] / sd[k][None, :]
        V[k] = Vk

    node(T, Q[T].copy())
    for k in reversed(range(T)):  # :660-720
        H = B[k].T @ W[k + 1]
        G = R[k] + H @ B[k]
        LG[k] = _chol(G)
        Fm = W[k + 1] @ A[k]
        H = M[k].T + B[k].T @ Fm
        K[k] = -_chol_solve(LG[k], H)
        node(k, Q[k] + A[k].T @ Fm + K[k].T @ H)

    finv = lambda k, rhs: sd[k] * _chol_solve(LF[k], rhs / sd[k])  # :531-549
    v, kk = [None] * (T + 1), [None] * T
    v[T] = vec(g["q"], T, n).copy()
    for k in reversed(range(T)):  # :738-796
        dc, cc = vec(g["delta"], k + 1, n), vec(g["c"], k + 1, n)
        f = dc * v[k + 1] - cc
        gg = v[k + 1] - W[k + 1] @ f
        h = vec(g["r"], k, m) + B[k].T @ gg
        kk[k] = -_chol_solve(LG[k], h)
        v[k] = vec(g["q"], k, n) + A[k].T @ gg + K[k].T @ h
    x, u, y = [None] * (T + 1), [None] * T, [None] * (T + 1)
    x[0] = -finv(0, vec(g["delta"], 0, n) * v[0] - vec(g["c"], 0, n))  # :798-819
    y[0] = v[0] + V[0] @ x[0]
    for k in range(T):  # :821-870
        u[k] = kk[k] + K[k] @ x[k]
        f = vec(g["c"], k + 1, n) - vec(g["delta"], k + 1, n) * v[k + 1] + A[k] @ x[k] + B[k] @ u[k]
        x[k + 1] = finv(k + 1, f)
        y[k + 1] = v[k + 1] + V[k + 1] @ x[k + 1]
    cat = lambda parts: np.concatenate(parts).astype(np.float64)
    return cat(x), cat(u), cat(y)


def lqr_chain_truth_batch(n, m, T, host):
    batch = host["q"].shape[0]
    outs = [lqr_chain_truth(n, m, T, {k: v[i] for k, v in host.items()}) for i in range(batch)]
    return dict(x=np.stack([o[0] for o in outs]), u=np.stack([o[1] for o in outs]),
                y=np.stack([o[2] for o in outs]))


def wide_delta(host, seed, lo=1e-3, hi=1e9):
    """delta log-uniform over the Newton-KKT benchmark's r2 range (newton_kkt_benchmark.cpp:231-233)."""
    rng = np.random.default_rng(seed)
    out = dict(host)
    out["delta"] = np.exp(rng.uniform(np.log(lo), np.log(hi), size=host["delta"].shape))
    return out


def error_against(truth, got):
    """max over (x, u, y) of the per-problem relative error ||got - truth|| / ||truth||."""
    worst = 0.0
    for k in ("x", "u", "y"):
        num = np.linalg.norm(got[k] - truth[k], axis=1)
        den = np.linalg.norm(truth[k], axis=1)
        worst = max(worst, float((num / den).max()))
    return worst
