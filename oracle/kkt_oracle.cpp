// ORACLE — TEST INFRASTRUCTURE ONLY (see kkt_oracle.hpp).
#include "kkt_oracle.hpp"

#include <algorithm>

namespace sipoc_oracle {

namespace {
int dim_or_zero(const int *dims, int i) { return dims == nullptr ? 0 : dims[i]; }
}  // namespace

// types.cpp:24-64 (populate_workspace_metadata) + lqr.cpp:146-180.
KktLayout make_kkt_layout(const Tree &tree, const int *state_dims,
                          const int *control_dims, const ConstraintDims &cd) {
  KktLayout K;
  K.lqr = make_layout(tree, state_dims, control_dims);
  const int E = tree.num_edges, N = E + 1;
  K.node_c.resize(N);
  K.node_g.resize(N);
  K.edge_c.resize(E);
  K.edge_g.resize(E);
  for (int i = 0; i < N; ++i) {
    K.node_c[i] = dim_or_zero(cd.node_c, i);
    K.node_g[i] = dim_or_zero(cd.node_g, i);
  }
  for (int e = 0; e < E; ++e) {
    K.edge_c[e] = dim_or_zero(cd.edge_c, e);
    K.edge_g[e] = dim_or_zero(cd.edge_g, e);
  }

  K.x_state.assign(N, 0);
  K.x_control.assign(E, 0);
  int off = 0;
  for (int node = 0; node < N; ++node) {  // types.cpp:33-41
    K.x_state[node] = off;
    if (node < E) {
      off += state_dims[node];
      K.x_control[node] = off;
      off += control_dims[node];
    }
  }
  // lqr.cpp:146-151: x_dim = n_E + sum_{e<E} (n_e + m_e)
  K.x_dim = state_dims[E];
  for (int e = 0; e < E; ++e) K.x_dim += state_dims[e] + control_dims[e];

  K.y_dyn.assign(N, 0);
  K.y_node_c.assign(N, 0);
  K.y_edge_c.assign(E, 0);
  off = 0;
  for (int node = 0; node < N; ++node) {  // types.cpp:43-49
    K.y_dyn[node] = off;
    off += state_dims[node];
    K.y_node_c[node] = off;
    off += K.node_c[node];
  }
  for (int e = 0; e < E; ++e) {  // types.cpp:50-53
    K.y_edge_c[e] = off;
    off += K.edge_c[e];
  }
  K.y_dim = off;

  K.z_node.assign(N, 0);
  K.z_edge.assign(E, 0);
  off = 0;
  for (int node = 0; node < N; ++node) {  // types.cpp:55-59
    K.z_node[node] = off;
    off += K.node_g[node];
  }
  for (int e = 0; e < E; ++e) {  // types.cpp:60-63
    K.z_edge[e] = off;
    off += K.edge_g[e];
  }
  K.z_dim = off;
  K.kkt_dim = K.x_dim + K.y_dim + K.z_dim;

  K.jc_node_off.assign(N + 1, 0);
  K.jg_node_off.assign(N + 1, 0);
  for (int i = 0; i < N; ++i) {
    K.jc_node_off[i + 1] = K.jc_node_off[i] + K.node_c[i] * state_dims[i];
    K.jg_node_off[i + 1] = K.jg_node_off[i] + K.node_g[i] * state_dims[i];
  }
  K.jcx_off.assign(E + 1, 0);
  K.jcu_off.assign(E + 1, 0);
  K.jgx_off.assign(E + 1, 0);
  K.jgu_off.assign(E + 1, 0);
  for (int e = 0; e < E; ++e) {
    const int np = state_dims[tree.parents[e]], m = control_dims[e];
    K.jcx_off[e + 1] = K.jcx_off[e] + K.edge_c[e] * np;
    K.jcu_off[e + 1] = K.jcu_off[e] + K.edge_c[e] * m;
    K.jgx_off[e + 1] = K.jgx_off[e] + K.edge_g[e] * np;
    K.jgu_off[e + 1] = K.jgu_off[e] + K.edge_g[e] * m;
  }
  return K;
}

void KktWorkspace::reserve(const KktLayout &K, const CompiledTree &tree) {
  const FlatLayout &L = K.lqr;
  const int E = L.num_edges, N = E + 1;
  hxx_edge_off.assign(E + 1, 0);
  for (int e = 0; e < E; ++e) {
    const int np = L.n[tree.parents[e]];
    hxx_edge_off[e + 1] = hxx_edge_off[e] + np * np;
  }
  Q_mod.assign(L.nn_off[N], 0.0);
  M_mod.assign(L.nm_off[E], 0.0);
  R_mod.assign(L.mm_off[E], 0.0);
  q_mod.assign(L.n_off[N], 0.0);
  r_mod.assign(L.m_off[E], 0.0);
  c_mod.assign(L.n_off[N], 0.0);
  dyn_r2.assign(L.n_off[N], 0.0);
  node_c_off.assign(N + 1, 0);
  node_g_off.assign(N + 1, 0);
  for (int i = 0; i < N; ++i) {
    node_c_off[i + 1] = node_c_off[i] + K.node_c[i];
    node_g_off[i + 1] = node_g_off[i] + K.node_g[i];
  }
  edge_c_off.assign(E + 1, 0);
  edge_g_off.assign(E + 1, 0);
  for (int e = 0; e < E; ++e) {
    edge_c_off[e + 1] = edge_c_off[e] + K.edge_c[e];
    edge_g_off[e + 1] = edge_g_off[e] + K.edge_g[e];
  }
  node_c_r2_inv.assign(node_c_off[N], 0.0);
  node_mod_w_inv.assign(node_g_off[N], 0.0);
  edge_c_r2_inv.assign(edge_c_off[E], 0.0);
  edge_mod_w_inv.assign(edge_g_off[E], 0.0);
  x.assign(L.n_off[N], 0.0);
  y.assign(L.n_off[N], 0.0);
  u.assign(L.m_off[E], 0.0);
  lqr.reserve(L, tree);
}

namespace {

// helpers.cpp:117-136: lower triangle of Q += J^T diag(w) J, one constraint
// row at a time, skipping exact zeros.
void add_weighted_state_jacobian_product(double *Q, int n, const double *J,
                                         int rows, const double *weights) {
  for (int k = 0; k < rows; ++k) {
    const double weight = weights[k];
    for (int col = 0; col < n; ++col) {
      const double wj = weight * J[k + col * rows];
      if (wj == 0.0) continue;
      for (int row = col; row < n; ++row) {
        const double j = J[k + row * rows];
        if (j == 0.0) continue;
        Q[row + col * n] += wj * j;
      }
    }
  }
}

// helpers.cpp:79-115: M += Jx^T diag(w) Ju (full), lower triangle of
// R += Ju^T diag(w) Ju.
void add_weighted_control_jacobian_products(double *M, double *R, int n, int m,
                                            const double *Jx, const double *Ju,
                                            int rows, const double *weights) {
  for (int k = 0; k < rows; ++k) {
    const double weight = weights[k];
    for (int col = 0; col < m; ++col) {
      const double wju = weight * Ju[k + col * rows];
      if (wju == 0.0) continue;
      for (int row = 0; row < n; ++row) {
        const double jx = Jx[k + row * rows];
        if (jx == 0.0) continue;
        M[row + col * n] += jx * wju;
      }
    }
    for (int col = 0; col < m; ++col) {
      const double wju = weight * Ju[k + col * rows];
      if (wju == 0.0) continue;
      for (int row = col; row < m; ++row) {
        const double ju = Ju[k + row * rows];
        if (ju == 0.0) continue;
        R[row + col * m] += wju * ju;
      }
    }
  }
}

// helpers.cpp:138-153: result -= J^T (w .* rhs).
void subtract_weighted_jacobian_rhs(double *result, int cols, const double *J,
                                    int rows, const double *weights,
                                    const double *rhs) {
  for (int k = 0; k < rows; ++k) {
    const double wr = weights[k] * rhs[k];
    for (int col = 0; col < cols; ++col) {
      const double j = J[k + col * rows];
      if (j == 0.0) continue;
      result[col] -= j * wr;
    }
  }
}

// helpers.cpp:155-158
void mirror_lower_to_upper(double *A, int n) {
  for (int col = 0; col < n; ++col)
    for (int row = col + 1; row < n; ++row) A[col + row * n] = A[row + col * n];
}

// y (rows) += J (rows x cols) x
void add_Jx(double *y, const double *J, int rows, int cols, const double *x) {
  for (int i = 0; i < rows; ++i) {
    double s = 0.0;
    for (int j = 0; j < cols; ++j) s += J[i + j * rows] * x[j];
    y[i] += s;
  }
}
// y (cols) += J^T (rows x cols) x
void add_JTx(double *y, const double *J, int rows, int cols, const double *x) {
  for (int j = 0; j < cols; ++j) {
    double s = 0.0;
    for (int i = 0; i < rows; ++i) s += J[i + j * rows] * x[i];
    y[j] += s;
  }
}

}  // namespace

// helpers.cpp:242-370
bool kkt_factor(const CompiledTree &tree, const KktLayout &K,
                const KktModel &mdl, const double *w, const double *r1,
                const double *r2, const double *r3, KktWorkspace &ws,
                int *lqr_status) {
  const FlatLayout &L = K.lqr;
  const int E = L.num_edges, N = E + 1;
  if (lqr_status) *lqr_status = -1;

  for (int node = 0; node < N; ++node) {  // :251-277
    for (int row = 0; row < L.n[node]; ++row) {
      const double reg = r2[K.y_dyn[node] + row];
      if (reg <= 0.0) return false;
      ws.dyn_r2[L.n_off[node] + row] = reg;
    }
    for (int row = 0; row < K.node_c[node]; ++row) {
      const double reg = r2[K.y_node_c[node] + row];
      if (reg <= 0.0) return false;
      ws.node_c_r2_inv[ws.node_c_off[node] + row] = 1.0 / reg;
    }
    for (int row = 0; row < K.node_g[node]; ++row) {
      const int o = K.z_node[node] + row;
      const double reg = w[o] + r3[o];
      if (reg <= 0.0) return false;
      ws.node_mod_w_inv[ws.node_g_off[node] + row] = 1.0 / reg;
    }
  }
  for (int e = 0; e < E; ++e) {  // :279-295
    for (int row = 0; row < K.edge_c[e]; ++row) {
      const double reg = r2[K.y_edge_c[e] + row];
      if (reg <= 0.0) return false;
      ws.edge_c_r2_inv[ws.edge_c_off[e] + row] = 1.0 / reg;
    }
    for (int row = 0; row < K.edge_g[e]; ++row) {
      const int o = K.z_edge[e] + row;
      const double reg = w[o] + r3[o];
      if (reg <= 0.0) return false;
      ws.edge_mod_w_inv[ws.edge_g_off[e] + row] = 1.0 / reg;
    }
  }

  for (int node = 0; node < N; ++node) {  // :297-316
    const int n = L.n[node];
    double *Q = ws.Q_mod.data() + L.nn_off[node];
    const double *H = mdl.node_hxx + L.nn_off[node];
    for (int col = 0; col < n; ++col)
      for (int row = 0; row < n; ++row)
        Q[row + col * n] = row >= col ? H[row + col * n] : 0.0;
    for (int i = 0; i < n; ++i) Q[i + i * n] += r1[K.x_state[node] + i];
    add_weighted_state_jacobian_product(Q, n, mdl.node_jc + K.jc_node_off[node],
                                        K.node_c[node],
                                        ws.node_c_r2_inv.data() + ws.node_c_off[node]);
    add_weighted_state_jacobian_product(Q, n, mdl.node_jg + K.jg_node_off[node],
                                        K.node_g[node],
                                        ws.node_mod_w_inv.data() + ws.node_g_off[node]);
  }

  for (int e = 0; e < E; ++e) {  // :318-354
    const int parent = tree.parents[e];
    const int n = L.n[parent], m = L.m[e];
    const int c = K.edge_c[e], g = K.edge_g[e];
    const double *Hxx = mdl.edge_hxx + ws.hxx_edge_off[e];
    const double *Hxu = mdl.edge_hxu + L.nm_off[e];
    const double *Huu = mdl.edge_huu + L.mm_off[e];
    const double *Jcx = mdl.edge_jcx + K.jcx_off[e];
    const double *Jcu = mdl.edge_jcu + K.jcu_off[e];
    const double *Jgx = mdl.edge_jgx + K.jgx_off[e];
    const double *Jgu = mdl.edge_jgu + K.jgu_off[e];
    const double *wc = ws.edge_c_r2_inv.data() + ws.edge_c_off[e];
    const double *wg = ws.edge_mod_w_inv.data() + ws.edge_g_off[e];

    double *Q = ws.Q_mod.data() + L.nn_off[parent];
    for (int col = 0; col < n; ++col)
      for (int row = col; row < n; ++row) Q[row + col * n] += Hxx[row + col * n];
    add_weighted_state_jacobian_product(Q, n, Jcx, c, wc);
    add_weighted_state_jacobian_product(Q, n, Jgx, g, wg);

    double *M = ws.M_mod.data() + L.nm_off[e];
    double *R = ws.R_mod.data() + L.mm_off[e];
    std::copy_n(Hxu, n * m, M);
    for (int col = 0; col < m; ++col)
      for (int row = 0; row < m; ++row)
        R[row + col * m] = row >= col ? Huu[row + col * m] : 0.0;
    for (int i = 0; i < m; ++i) R[i + i * m] += r1[K.x_control[e] + i];
    add_weighted_control_jacobian_products(M, R, n, m, Jcx, Jcu, c, wc);
    add_weighted_control_jacobian_products(M, R, n, m, Jgx, Jgu, g, wg);
    mirror_lower_to_upper(R, m);
  }

  for (int node = 0; node < N; ++node)  // :356-360
    mirror_lower_to_upper(ws.Q_mod.data() + L.nn_off[node], L.n[node]);

  // :362-368
  LqrInput in{ws.Q_mod.data(), ws.M_mod.data(), ws.R_mod.data(), nullptr, nullptr,
              mdl.edge_A, mdl.edge_B, nullptr, ws.dyn_r2.data()};
  const Status st = lqr_factor(tree, L, in, ws.lqr);
  if (lqr_status) *lqr_status = static_cast<int>(st);
  return st == SUCCESS;
}

// helpers.cpp:749-894
void kkt_solve(const CompiledTree &tree, const KktLayout &K, const KktModel &mdl,
               const double *b, double *sol, KktWorkspace &ws) {
  const FlatLayout &L = K.lqr;
  const int E = L.num_edges, N = E + 1;
  const int x_dim = K.x_dim, y_dim = K.y_dim;

  for (int node = 0; node < N; ++node) {  // :752-778
    const int n = L.n[node];
    double *q = ws.q_mod.data() + L.n_off[node];
    double *cm = ws.c_mod.data() + L.n_off[node];
    for (int i = 0; i < n; ++i) q[i] = -b[K.x_state[node] + i];
    subtract_weighted_jacobian_rhs(q, n, mdl.node_jc + K.jc_node_off[node],
                                   K.node_c[node],
                                   ws.node_c_r2_inv.data() + ws.node_c_off[node],
                                   b + x_dim + K.y_node_c[node]);
    subtract_weighted_jacobian_rhs(q, n, mdl.node_jg + K.jg_node_off[node],
                                   K.node_g[node],
                                   ws.node_mod_w_inv.data() + ws.node_g_off[node],
                                   b + x_dim + y_dim + K.z_node[node]);
    for (int i = 0; i < n; ++i) cm[i] = -b[x_dim + K.y_dyn[node] + i];
  }

  for (int e = 0; e < E; ++e) {  // :780-812
    const int parent = tree.parents[e];
    const int np = L.n[parent], m = L.m[e];
    const int c = K.edge_c[e], g = K.edge_g[e];
    const double *wc = ws.edge_c_r2_inv.data() + ws.edge_c_off[e];
    const double *wg = ws.edge_mod_w_inv.data() + ws.edge_g_off[e];
    const double *b_yc = b + x_dim + K.y_edge_c[e];
    const double *b_z = b + x_dim + y_dim + K.z_edge[e];
    double *qp = ws.q_mod.data() + L.n_off[parent];
    double *r = ws.r_mod.data() + L.m_off[e];
    subtract_weighted_jacobian_rhs(qp, np, mdl.edge_jcx + K.jcx_off[e], c, wc, b_yc);
    subtract_weighted_jacobian_rhs(qp, np, mdl.edge_jgx + K.jgx_off[e], g, wg, b_z);
    for (int i = 0; i < m; ++i) r[i] = -b[K.x_control[e] + i];
    subtract_weighted_jacobian_rhs(r, m, mdl.edge_jcu + K.jcu_off[e], c, wc, b_yc);
    subtract_weighted_jacobian_rhs(r, m, mdl.edge_jgu + K.jgu_off[e], g, wg, b_z);
  }

  // :814-826 (the reference aims Output pointers into sol; we scatter after).
  LqrInput in{ws.Q_mod.data(), ws.M_mod.data(), ws.R_mod.data(), ws.q_mod.data(),
              ws.r_mod.data(), mdl.edge_A, mdl.edge_B, ws.c_mod.data(),
              ws.dyn_r2.data()};
  LqrOutput out{ws.x.data(), ws.u.data(), ws.y.data()};
  lqr_solve(tree, L, in, ws.lqr, out);
  for (int node = 0; node < N; ++node)
    for (int i = 0; i < L.n[node]; ++i) {
      sol[K.x_state[node] + i] = ws.x[L.n_off[node] + i];
      sol[x_dim + K.y_dyn[node] + i] = ws.y[L.n_off[node] + i];
    }
  for (int e = 0; e < E; ++e)
    for (int i = 0; i < L.m[e]; ++i) sol[K.x_control[e] + i] = ws.u[L.m_off[e] + i];

  for (int node = 0; node < N; ++node) {  // :828-856
    const int n = L.n[node], c = K.node_c[node], g = K.node_g[node];
    const double *x = sol + K.x_state[node];
    double *y_c = sol + x_dim + K.y_node_c[node];
    double *z = sol + x_dim + y_dim + K.z_node[node];
    for (int i = 0; i < c; ++i) y_c[i] = 0.0;
    for (int i = 0; i < g; ++i) z[i] = 0.0;
    add_Jx(y_c, mdl.node_jc + K.jc_node_off[node], c, n, x);
    add_Jx(z, mdl.node_jg + K.jg_node_off[node], g, n, x);
    for (int i = 0; i < c; ++i) {
      y_c[i] -= b[x_dim + K.y_node_c[node] + i];
      y_c[i] = ws.node_c_r2_inv[ws.node_c_off[node] + i] * y_c[i];
    }
    for (int i = 0; i < g; ++i) {
      z[i] -= b[x_dim + y_dim + K.z_node[node] + i];
      z[i] = ws.node_mod_w_inv[ws.node_g_off[node] + i] * z[i];
    }
  }

  for (int e = 0; e < E; ++e) {  // :858-893
    const int parent = tree.parents[e];
    const int n = L.n[parent], m = L.m[e];
    const int c = K.edge_c[e], g = K.edge_g[e];
    const double *xp = sol + K.x_state[parent];
    const double *u = sol + K.x_control[e];
    double *y_c = sol + x_dim + K.y_edge_c[e];
    double *z = sol + x_dim + y_dim + K.z_edge[e];
    for (int i = 0; i < c; ++i) y_c[i] = 0.0;
    add_Jx(y_c, mdl.edge_jcx + K.jcx_off[e], c, n, xp);
    add_Jx(y_c, mdl.edge_jcu + K.jcu_off[e], c, m, u);
    for (int i = 0; i < c; ++i) {
      y_c[i] -= b[x_dim + K.y_edge_c[e] + i];
      y_c[i] = ws.edge_c_r2_inv[ws.edge_c_off[e] + i] * y_c[i];
    }
    for (int i = 0; i < g; ++i) z[i] = 0.0;
    add_Jx(z, mdl.edge_jgx + K.jgx_off[e], g, n, xp);
    add_Jx(z, mdl.edge_jgu + K.jgu_off[e], g, m, u);
    for (int i = 0; i < g; ++i) {
      z[i] -= b[x_dim + y_dim + K.z_edge[e] + i];
      z[i] = ws.edge_mod_w_inv[ws.edge_g_off[e] + i] * z[i];
    }
  }
}

// helpers.cpp:953-977 with the theta == 0 branches of add_{H,C,CT,G,GT}x_to_y
// (:979-1368).
void kkt_apply(const CompiledTree &tree, const KktLayout &K, const KktModel &mdl,
               const double *w, const double *r1, const double *r2,
               const double *r3, const double *x_x, const double *x_y,
               const double *x_z, double *y_x, double *y_y, double *y_z) {
  const FlatLayout &L = K.lqr;
  const int E = L.num_edges, N = E + 1;
  std::vector<int> hxx_edge_off(E + 1, 0);
  for (int e = 0; e < E; ++e) {
    const int np = L.n[tree.parents[e]];
    hxx_edge_off[e + 1] = hxx_edge_off[e] + np * np;
  }

  // add_Hx_to_y (:979-1017): full (not lower-only) Hessian blocks.
  for (int node = 0; node < N; ++node)
    add_Jx(y_x + K.x_state[node], mdl.node_hxx + L.nn_off[node], L.n[node],
           L.n[node], x_x + K.x_state[node]);
  for (int e = 0; e < E; ++e) {
    const int parent = tree.parents[e];
    const int n = L.n[parent], m = L.m[e];
    const double *xp = x_x + K.x_state[parent];
    const double *u = x_x + K.x_control[e];
    add_Jx(y_x + K.x_state[parent], mdl.edge_hxx + hxx_edge_off[e], n, n, xp);
    add_Jx(y_x + K.x_state[parent], mdl.edge_hxu + L.nm_off[e], n, m, u);
    add_JTx(y_x + K.x_control[e], mdl.edge_hxu + L.nm_off[e], n, m, xp);
    add_Jx(y_x + K.x_control[e], mdl.edge_huu + L.mm_off[e], m, m, u);
  }

  // add_Cx_to_y (:1070-1125)
  {
    const int root = tree.preorder[0];
    for (int i = 0; i < L.n[root]; ++i)
      y_y[K.y_dyn[root] + i] -= x_x[K.x_state[root] + i];
  }
  for (int node = 0; node < N; ++node)
    add_Jx(y_y + K.y_node_c[node], mdl.node_jc + K.jc_node_off[node],
           K.node_c[node], L.n[node], x_x + K.x_state[node]);
  for (int e = 0; e < E; ++e) {
    const int parent = tree.parents[e], child = tree.children[e];
    const int np = L.n[parent], nc = L.n[child], m = L.m[e], c = K.edge_c[e];
    const double *xp = x_x + K.x_state[parent];
    const double *u = x_x + K.x_control[e];
    double *yd = y_y + K.y_dyn[child];
    add_Jx(yd, mdl.edge_A + L.a_off[e], nc, np, xp);
    add_Jx(yd, mdl.edge_B + L.b_off[e], nc, m, u);
    for (int i = 0; i < nc; ++i) yd[i] -= x_x[K.x_state[child] + i];
    add_Jx(y_y + K.y_edge_c[e], mdl.edge_jcx + K.jcx_off[e], c, np, xp);
    add_Jx(y_y + K.y_edge_c[e], mdl.edge_jcu + K.jcu_off[e], c, m, u);
  }

  // add_CTx_to_y (:1161-1216)
  {
    const int root = tree.preorder[0];
    for (int i = 0; i < L.n[root]; ++i)
      y_x[K.x_state[root] + i] -= x_y[K.y_dyn[root] + i];
  }
  for (int node = 0; node < N; ++node)
    add_JTx(y_x + K.x_state[node], mdl.node_jc + K.jc_node_off[node],
            K.node_c[node], L.n[node], x_y + K.y_node_c[node]);
  for (int e = 0; e < E; ++e) {
    const int parent = tree.parents[e], child = tree.children[e];
    const int np = L.n[parent], nc = L.n[child], m = L.m[e], c = K.edge_c[e];
    const double *dyn = x_y + K.y_dyn[child];
    const double *cv = x_y + K.y_edge_c[e];
    add_JTx(y_x + K.x_state[parent], mdl.edge_A + L.a_off[e], nc, np, dyn);
    add_JTx(y_x + K.x_state[parent], mdl.edge_jcx + K.jcx_off[e], c, np, cv);
    for (int i = 0; i < nc; ++i) y_x[K.x_state[child] + i] -= dyn[i];
    add_JTx(y_x + K.x_control[e], mdl.edge_B + L.b_off[e], nc, m, dyn);
    add_JTx(y_x + K.x_control[e], mdl.edge_jcu + K.jcu_off[e], c, m, cv);
  }

  // add_Gx_to_y (:1252-1284)
  for (int node = 0; node < N; ++node)
    add_Jx(y_z + K.z_node[node], mdl.node_jg + K.jg_node_off[node],
           K.node_g[node], L.n[node], x_x + K.x_state[node]);
  for (int e = 0; e < E; ++e) {
    const int parent = tree.parents[e];
    const int np = L.n[parent], m = L.m[e], g = K.edge_g[e];
    add_Jx(y_z + K.z_edge[e], mdl.edge_jgx + K.jgx_off[e], g, np,
           x_x + K.x_state[parent]);
    add_Jx(y_z + K.z_edge[e], mdl.edge_jgu + K.jgu_off[e], g, m,
           x_x + K.x_control[e]);
  }

  // add_GTx_to_y (:1311-1343)
  for (int node = 0; node < N; ++node)
    add_JTx(y_x + K.x_state[node], mdl.node_jg + K.jg_node_off[node],
            K.node_g[node], L.n[node], x_z + K.z_node[node]);
  for (int e = 0; e < E; ++e) {
    const int parent = tree.parents[e];
    const int np = L.n[parent], m = L.m[e], g = K.edge_g[e];
    add_JTx(y_x + K.x_state[parent], mdl.edge_jgx + K.jgx_off[e], g, np,
            x_z + K.z_edge[e]);
    add_JTx(y_x + K.x_control[e], mdl.edge_jgu + K.jgu_off[e], g, m,
            x_z + K.z_edge[e]);
  }

  // :968-976
  for (int i = 0; i < K.x_dim; ++i) y_x[i] += r1[i] * x_x[i];
  for (int i = 0; i < K.y_dim; ++i) y_y[i] -= r2[i] * x_y[i];
  for (int i = 0; i < K.z_dim; ++i) y_z[i] -= (w[i] + r3[i]) * x_z[i];
}

}  // namespace sipoc_oracle
