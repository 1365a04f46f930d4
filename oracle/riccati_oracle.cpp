// ORACLE — TEST INFRASTRUCTURE ONLY (see riccati_oracle.hpp).
//
// Restates sip_optimal_control/lqr.cpp of the reference with plain loops over
// column-major arrays.  Each function cites the reference lines it follows.
#include "riccati_oracle.hpp"

#include <algorithm>
#include <cmath>

namespace sipoc_oracle {

// ---------------------------------------------------------------------------
// Topology: lqr.cpp:563-631 (compile_topology_data).
// ---------------------------------------------------------------------------
CompiledTree compile_tree(const Tree &tree) {
  CompiledTree out;
  const int E = tree.num_edges;
  const int N = tree.num_nodes();
  out.status = INVALID_TOPOLOGY;
  if (E < 0) return out;
  if (E > 0 && (tree.parents == nullptr || tree.children == nullptr)) return out;
  if (tree.root < 0 || tree.root >= N) return out;  // lqr.cpp:571-574

  out.child_offsets.assign(N + 1, 0);
  out.child_edges.assign(E, 0);
  out.parents.assign(E, 0);
  out.children.assign(E, 0);
  out.preorder.assign(N, 0);
  out.postorder.assign(N, 0);

  for (int e = 0; e < E; ++e) {  // lqr.cpp:577-587
    const int p = tree.parents[e], c = tree.children[e];
    if (p < 0 || p >= N || c < 0 || c >= N || p == c) return out;
    out.parents[e] = p;
    out.children[e] = c;
    ++out.child_offsets[p + 1];
  }
  for (int i = 0; i < N; ++i) out.child_offsets[i + 1] += out.child_offsets[i];

  // Fill the CSR child lists in edge order (lqr.cpp:593-598).
  std::vector<int> cursor(out.child_offsets.begin(), out.child_offsets.end() - 1);
  for (int e = 0; e < E; ++e) out.child_edges[cursor[out.parents[e]]++] = e;

  // Iterative DFS; children pushed in reverse so the first child is visited
  // first (lqr.cpp:600-619).  A node reached twice, or more than N visits,
  // means the graph is not a tree.
  std::vector<int> stack;
  stack.reserve(N + E);
  std::vector<int> mark(N, 0);
  stack.push_back(tree.root);
  int visited = 0;
  while (!stack.empty()) {
    const int node = stack.back();
    stack.pop_back();
    if (visited >= N || mark[node] != 0) return out;
    mark[node] = 1;
    out.preorder[visited++] = node;
    for (int ci = out.child_offsets[node + 1] - 1; ci >= out.child_offsets[node]; --ci)
      stack.push_back(out.children[out.child_edges[ci]]);
  }
  if (visited != N) return out;  // disconnected (lqr.cpp:621-623)

  for (int i = 0; i < N; ++i) out.postorder[i] = out.preorder[N - 1 - i];  // :625-628
  out.status = SUCCESS;
  return out;
}

FlatLayout make_layout(const Tree &tree, const int *state_dims,
                       const int *control_dims) {
  FlatLayout L;
  const int E = tree.num_edges, N = E + 1;
  L.num_edges = E;
  L.n.assign(state_dims, state_dims + N);
  L.m.assign(control_dims, control_dims + E);
  L.nn_off.assign(N + 1, 0);
  L.n_off.assign(N + 1, 0);
  for (int i = 0; i < N; ++i) {
    L.nn_off[i + 1] = L.nn_off[i] + L.n[i] * L.n[i];
    L.n_off[i + 1] = L.n_off[i] + L.n[i];
  }
  L.nm_off.assign(E + 1, 0);
  L.mm_off.assign(E + 1, 0);
  L.m_off.assign(E + 1, 0);
  L.a_off.assign(E + 1, 0);
  L.b_off.assign(E + 1, 0);
  for (int e = 0; e < E; ++e) {
    const int np = L.n[tree.parents[e]], nc = L.n[tree.children[e]], m = L.m[e];
    L.nm_off[e + 1] = L.nm_off[e] + np * m;
    L.mm_off[e + 1] = L.mm_off[e] + m * m;
    L.m_off[e + 1] = L.m_off[e] + m;
    L.a_off[e + 1] = L.a_off[e] + nc * np;
    L.b_off[e + 1] = L.b_off[e] + nc * m;
  }
  return L;
}

void LqrWorkspace::reserve(const FlatLayout &L, const CompiledTree &tree) {
  const int E = L.num_edges, N = E + 1;
  w_off.assign(E + 1, 0);
  k_off.assign(E + 1, 0);
  int max_n = 0, max_m = 0;
  for (int i = 0; i < N; ++i) max_n = std::max(max_n, L.n[i]);
  for (int e = 0; e < E; ++e) {
    max_m = std::max(max_m, L.m[e]);
    const int nc = L.n[tree.children[e]], np = L.n[tree.parents[e]];
    w_off[e + 1] = w_off[e] + nc * nc;
    k_off[e + 1] = k_off[e] + L.m[e] * np;
  }
  W.assign(w_off[E], 0.0);
  K.assign(k_off[E], 0.0);
  G_factor.assign(L.mm_off[E], 0.0);
  k.assign(L.m_off[E], 0.0);
  V.assign(L.nn_off[N], 0.0);
  F_factor.assign(L.nn_off[N], 0.0);
  sqrt_delta.assign(L.n_off[N], 0.0);
  sqrt_delta_inv.assign(L.n_off[N], 0.0);
  v.assign(L.n_off[N], 0.0);
  H.assign(std::max(1, max_m * max_n), 0.0);
  F.assign(std::max(1, max_n * max_n), 0.0);
  f.assign(std::max(1, max_n), 0.0);
  g.assign(std::max(1, max_n), 0.0);
  h.assign(std::max(1, max_m), 0.0);
}

// ---------------------------------------------------------------------------
// Dense helpers (column-major, ld = rows).
// ---------------------------------------------------------------------------

// Eigen 3.4.0 llt_inplace<Scalar, Lower>::unblocked: row-by-row left-looking
// update, pivot test x <= 0, column scaled by division.
bool cholesky_lower_inplace(double *a, const int n) {
  for (int k = 0; k < n; ++k) {
    double x = a[k + k * n];
    for (int j = 0; j < k; ++j) x -= a[k + j * n] * a[k + j * n];
    if (x <= 0.0) return false;
    x = std::sqrt(x);
    a[k + k * n] = x;
    for (int i = k + 1; i < n; ++i) {
      double s = a[i + k * n];
      for (int j = 0; j < k; ++j) s -= a[i + j * n] * a[k + j * n];
      a[i + k * n] = s / x;
    }
  }
  return true;
}

// triangularView<Lower>().solveInPlace followed by
// transpose().triangularView<Upper>().solveInPlace, one right-hand side
// column at a time (lqr.cpp:517-519, 542-544, 708-712, 786-790).
void cholesky_solve_inplace(const double *l, const int n, double *b,
                            const int nrhs) {
  for (int c = 0; c < nrhs; ++c) {
    double *x = b + c * n;
    for (int i = 0; i < n; ++i) {
      double s = x[i];
      for (int j = 0; j < i; ++j) s -= l[i + j * n] * x[j];
      x[i] = s / l[i + i * n];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = x[i];
      for (int j = i + 1; j < n; ++j) s -= l[j + i * n] * x[j];
      x[i] = s / l[i + i * n];
    }
  }
}

namespace {

// lqr.cpp:475-485
bool compute_delta_sqrt(const double *delta, double *sd, double *sdi, int n) {
  for (int i = 0; i < n; ++i) {
    if (delta[i] <= 0.0) return false;
    sd[i] = std::sqrt(delta[i]);
    sdi[i] = 1.0 / sd[i];
  }
  return true;
}

// lqr.cpp:487-509: F = I + D^1/2 V D^1/2 (all n*n entries), lower LLT.
Status factor_F(const double *delta, const double *V, double *Ff, double *sd,
                double *sdi, int n) {
  if (!compute_delta_sqrt(delta, sd, sdi, n)) return INVALID_DELTA;
  for (int col = 0; col < n; ++col) {
    for (int row = 0; row < n; ++row)
      Ff[row + col * n] = sd[row] * V[row + col * n] * sd[col];
    Ff[col + col * n] += 1.0;
  }
  return cholesky_lower_inplace(Ff, n) ? SUCCESS : F_FACTORIZATION_FAILURE;
}

// lqr.cpp:511-529: W = D^-1/2 (I - F^-1) D^-1/2 via two triangular solves on
// the identity.
void compute_regularized_W(const double *Ff, double *W, const double *sdi, int n) {
  for (int i = 0; i < n * n; ++i) W[i] = 0.0;
  for (int i = 0; i < n; ++i) W[i + i * n] = 1.0;
  cholesky_solve_inplace(Ff, n, W, n);
  for (int i = 0; i < n * n; ++i) W[i] *= -1.0;
  for (int i = 0; i < n; ++i) W[i + i * n] += 1.0;
  for (int col = 0; col < n; ++col)
    for (int row = 0; row < n; ++row) W[row + col * n] *= sdi[row] * sdi[col];
}

// lqr.cpp:531-549: (I + D V)^-1 rhs = D^1/2 F^-1 D^-1/2 rhs.
void F_inv_mult_vector(const double *Ff, const double *rhs, double *result,
                       const double *sd, const double *sdi, int n) {
  for (int i = 0; i < n; ++i) result[i] = sdi[i] * rhs[i];
  cholesky_solve_inplace(Ff, n, result, 1);
  for (int i = 0; i < n; ++i) result[i] *= sd[i];
}

// C (r x c) (+)= op(A) * op(B) helpers, column-major.
// C = A^T * B, A is (k x r), B is (k x c).
void gemm_tn(double *C, const double *A, const double *B, int r, int c, int k,
             bool accumulate) {
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) {
      double s = accumulate ? C[i + j * r] : 0.0;
      for (int p = 0; p < k; ++p) s += A[p + i * k] * B[p + j * k];
      C[i + j * r] = s;
    }
}
// C = A * B, A is (r x k), B is (k x c).
void gemm_nn(double *C, const double *A, const double *B, int r, int c, int k,
             bool accumulate) {
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) {
      double s = accumulate ? C[i + j * r] : 0.0;
      for (int p = 0; p < k; ++p) s += A[i + p * r] * B[p + j * k];
      C[i + j * r] = s;
    }
}

}  // namespace

// ---------------------------------------------------------------------------
// lqr.cpp:645-731 (LQR::factor_with_status)
// ---------------------------------------------------------------------------
Status lqr_factor(const CompiledTree &tree, const FlatLayout &L,
                  const LqrInput &in, LqrWorkspace &ws) {
  if (tree.status != SUCCESS) return tree.status;  // :646-648
  const int N = L.num_edges + 1;

  for (int order = 0; order < N; ++order) {
    const int node = tree.postorder[order];
    const int n = L.n[node];
    double *V = ws.V.data() + L.nn_off[node];
    std::copy_n(in.Q + L.nn_off[node], n * n, V);  // :658

    for (int ci = tree.child_offsets[node]; ci < tree.child_offsets[node + 1]; ++ci) {
      const int e = tree.child_edges[ci];
      const int child = tree.children[e];
      const int nc = L.n[child];
      const int m = L.m[e];
      const double *A = in.A + L.a_off[e];  // nc x n
      const double *B = in.B + L.b_off[e];  // nc x m
      const double *M = in.M + L.nm_off[e]; // n x m
      const double *R = in.R + L.mm_off[e]; // m x m
      double *W = ws.W.data() + ws.w_off[e];
      double *Gf = ws.G_factor.data() + L.mm_off[e];
      double *K = ws.K.data() + ws.k_off[e];  // m x n
      double *H = ws.H.data();
      double *Fcp = ws.F.data();

      compute_regularized_W(ws.F_factor.data() + L.nn_off[child], W,
                            ws.sqrt_delta_inv.data() + L.n_off[child], nc);  // :689

      gemm_tn(H, B, W, m, nc, nc, false);  // H_child = B^T W  (:692)
      std::copy_n(R, m * m, Gf);           // :693
      gemm_nn(Gf, H, B, m, m, nc, true);   // G += H_child B  (:694)
      if (!cholesky_lower_inplace(Gf, m)) return G_FACTORIZATION_FAILURE;  // :696-701

      gemm_nn(Fcp, W, A, nc, n, nc, false);  // F = W A  (:703)
      for (int j = 0; j < n; ++j)            // H_parent = M^T  (:704)
        for (int i = 0; i < m; ++i) H[i + j * m] = M[j + i * n];
      gemm_tn(H, B, Fcp, m, n, nc, true);    // += B^T F  (:705)

      std::copy_n(H, m * n, K);              // :707
      cholesky_solve_inplace(Gf, m, K, n);   // :708-712
      for (int i = 0; i < m * n; ++i) K[i] *= -1.0;  // :713

      gemm_tn(V, A, Fcp, n, n, nc, true);    // V += A^T F  (:715)
      // F_parent = K^T H ; V += F_parent  (:716-719)
      double *Fp = ws.F.data();
      gemm_tn(Fp, K, H, n, n, m, false);
      for (int i = 0; i < n * n; ++i) V[i] += Fp[i];
    }

    const Status st = factor_F(in.delta + L.n_off[node], V,
                               ws.F_factor.data() + L.nn_off[node],
                               ws.sqrt_delta.data() + L.n_off[node],
                               ws.sqrt_delta_inv.data() + L.n_off[node], n);  // :722-727
    if (st != SUCCESS) return st;
  }
  return SUCCESS;
}

// ---------------------------------------------------------------------------
// lqr.cpp:735-871 (LQR::solve)
// ---------------------------------------------------------------------------
void lqr_solve(const CompiledTree &tree, const FlatLayout &L, const LqrInput &in,
               LqrWorkspace &ws, const LqrOutput &out) {
  const int N = L.num_edges + 1;

  // Backward affine sweep (:738-796).
  for (int order = 0; order < N; ++order) {
    const int node = tree.postorder[order];
    const int n = L.n[node];
    double *v = ws.v.data() + L.n_off[node];
    std::copy_n(in.q + L.n_off[node], n, v);  // :744

    for (int ci = tree.child_offsets[node]; ci < tree.child_offsets[node + 1]; ++ci) {
      const int e = tree.child_edges[ci];
      const int child = tree.children[e];
      const int nc = L.n[child];
      const int m = L.m[e];
      const double *A = in.A + L.a_off[e];
      const double *B = in.B + L.b_off[e];
      const double *r = in.r + L.m_off[e];
      const double *cc = in.c + L.n_off[child];
      const double *dc = in.delta + L.n_off[child];
      const double *vc = ws.v.data() + L.n_off[child];
      const double *W = ws.W.data() + ws.w_off[e];
      const double *Gf = ws.G_factor.data() + L.mm_off[e];
      const double *K = ws.K.data() + ws.k_off[e];
      double *ke = ws.k.data() + L.m_off[e];
      double *f = ws.f.data(), *g = ws.g.data(), *h = ws.h.data();

      for (int i = 0; i < nc; ++i) f[i] = dc[i] * vc[i] - cc[i];  // :778-779
      for (int i = 0; i < nc; ++i) {                               // :780-781
        double s = 0.0;
        for (int j = 0; j < nc; ++j) s += W[i + j * nc] * f[j];
        g[i] = vc[i] - s;
      }
      for (int a = 0; a < m; ++a) {                                // :783-784
        double s = 0.0;
        for (int i = 0; i < nc; ++i) s += B[i + a * nc] * g[i];
        h[a] = r[a] + s;
      }
      std::copy_n(h, m, ke);                                       // :785
      cholesky_solve_inplace(Gf, m, ke, 1);                        // :786-790
      for (int a = 0; a < m; ++a) ke[a] *= -1.0;                   // :791

      for (int j = 0; j < n; ++j) {                                // :793-794
        double s = 0.0;
        for (int i = 0; i < nc; ++i) s += A[i + j * nc] * g[i];
        double t = 0.0;
        for (int a = 0; a < m; ++a) t += K[a + j * m] * h[a];
        v[j] += s;
        v[j] += t;
      }
    }
  }

  // Root (:798-819).
  {
    const int root = tree.preorder[0];
    const int n = L.n[root];
    const double *c = in.c + L.n_off[root];
    const double *d = in.delta + L.n_off[root];
    const double *V = ws.V.data() + L.nn_off[root];
    const double *v = ws.v.data() + L.n_off[root];
    double *x = out.x + L.n_off[root];
    double *y = out.y + L.n_off[root];
    double *f = ws.f.data();
    for (int i = 0; i < n; ++i) f[i] = d[i] * v[i] - c[i];
    F_inv_mult_vector(ws.F_factor.data() + L.nn_off[root], f, x,
                      ws.sqrt_delta.data() + L.n_off[root],
                      ws.sqrt_delta_inv.data() + L.n_off[root], n);
    for (int i = 0; i < n; ++i) x[i] *= -1.0;
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += V[i + j * n] * x[j];
      y[i] = v[i] + s;
    }
  }

  // Forward rollout (:821-870).
  for (int order = 0; order < N; ++order) {
    const int node = tree.preorder[order];
    const int n = L.n[node];
    const double *x = out.x + L.n_off[node];

    for (int ci = tree.child_offsets[node]; ci < tree.child_offsets[node + 1]; ++ci) {
      const int e = tree.child_edges[ci];
      const int child = tree.children[e];
      const int nc = L.n[child];
      const int m = L.m[e];
      const double *A = in.A + L.a_off[e];
      const double *B = in.B + L.b_off[e];
      const double *K = ws.K.data() + ws.k_off[e];
      const double *ke = ws.k.data() + L.m_off[e];
      const double *Vc = ws.V.data() + L.nn_off[child];
      const double *vc = ws.v.data() + L.n_off[child];
      const double *cc = in.c + L.n_off[child];
      const double *dc = in.delta + L.n_off[child];
      double *u = out.u + L.m_off[e];
      double *xc = out.x + L.n_off[child];
      double *yc = out.y + L.n_off[child];
      double *f = ws.f.data();

      for (int a = 0; a < m; ++a) {  // u = k + K x  (:856-857)
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += K[a + j * m] * x[j];
        u[a] = ke[a] + s;
      }
      for (int i = 0; i < nc; ++i) {  // :859-862
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += A[i + j * nc] * x[j];
        double t = 0.0;
        for (int a = 0; a < m; ++a) t += B[i + a * nc] * u[a];
        f[i] = cc[i] - dc[i] * vc[i];
        f[i] += s;
        f[i] += t;
      }
      F_inv_mult_vector(ws.F_factor.data() + L.nn_off[child], f, xc,
                        ws.sqrt_delta.data() + L.n_off[child],
                        ws.sqrt_delta_inv.data() + L.n_off[child], nc);  // :863-865
      for (int i = 0; i < nc; ++i) {  // y = v + V x  (:867-868)
        double s = 0.0;
        for (int j = 0; j < nc; ++j) s += Vc[i + j * nc] * xc[j];
        yc[i] = vc[i] + s;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// KKT residual of the LQR system.  tests/lqr_test.cpp:152-186 (chain) and
// :371-409 / :600-639 (tree); Q and R enter through their lower triangles
// (selfadjointView<Lower>).
// ---------------------------------------------------------------------------
double lqr_residual_norm(const CompiledTree &tree, const FlatLayout &L,
                         const LqrInput &in, const LqrOutput &out) {
  const int E = L.num_edges, N = E + 1;
  double sq = 0.0;
  std::vector<double> res;

  auto sym_lower_mv = [](const double *S, const double *x, int n, double *y) {
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int j = 0; j < n; ++j)
        s += (i >= j ? S[i + j * n] : S[j + i * n]) * x[j];
      y[i] += s;
    }
  };

  for (int node = 0; node < N; ++node) {
    const int n = L.n[node];
    res.assign(n, 0.0);
    sym_lower_mv(in.Q + L.nn_off[node], out.x + L.n_off[node], n, res.data());
    for (int i = 0; i < n; ++i)
      res[i] += in.q[L.n_off[node] + i] - out.y[L.n_off[node] + i];
    for (int ci = tree.child_offsets[node]; ci < tree.child_offsets[node + 1]; ++ci) {
      const int e = tree.child_edges[ci];
      const int child = tree.children[e];
      const int nc = L.n[child], m = L.m[e];
      const double *A = in.A + L.a_off[e];
      const double *M = in.M + L.nm_off[e];
      for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int a = 0; a < m; ++a) s += M[i + a * n] * out.u[L.m_off[e] + a];
        for (int p = 0; p < nc; ++p) s += A[p + i * nc] * out.y[L.n_off[child] + p];
        res[i] += s;
      }
    }
    for (int i = 0; i < n; ++i) sq += res[i] * res[i];
  }

  for (int e = 0; e < E; ++e) {
    const int parent = tree.parents[e], child = tree.children[e];
    const int n = L.n[parent], nc = L.n[child], m = L.m[e];
    const double *A = in.A + L.a_off[e];
    const double *B = in.B + L.b_off[e];
    const double *M = in.M + L.nm_off[e];
    const double *xp = out.x + L.n_off[parent];
    const double *xc = out.x + L.n_off[child];
    const double *yc = out.y + L.n_off[child];
    const double *u = out.u + L.m_off[e];
    res.assign(m, 0.0);
    sym_lower_mv(in.R + L.mm_off[e], u, m, res.data());
    for (int a = 0; a < m; ++a) {
      double s = in.r[L.m_off[e] + a];
      for (int i = 0; i < n; ++i) s += M[i + a * n] * xp[i];
      for (int p = 0; p < nc; ++p) s += B[p + a * nc] * yc[p];
      res[a] += s;
      sq += res[a] * res[a];
    }
    for (int p = 0; p < nc; ++p) {
      double s = in.c[L.n_off[child] + p] - xc[p] -
                 in.delta[L.n_off[child] + p] * yc[p];
      for (int i = 0; i < n; ++i) s += A[p + i * nc] * xp[i];
      for (int a = 0; a < m; ++a) s += B[p + a * nc] * u[a];
      sq += s * s;
    }
  }

  {
    const int root = tree.preorder[0];
    for (int i = 0; i < L.n[root]; ++i) {
      const int o = L.n_off[root] + i;
      const double s = -out.x[o] - in.delta[o] * out.y[o] + in.c[o];
      sq += s * s;
    }
  }
  return std::sqrt(sq);
}

}  // namespace sipoc_oracle
