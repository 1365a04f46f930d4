// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C entry points over the Eigen-free restatement (riccati_oracle.cpp,
// kkt_oracle.cpp) so tests/ and bench.py's cpu_baseline can drive it through
// ctypes.  Batches are problem-major: array X holds [batch][flat per-problem
// index]; problems are independent and run one per OpenMP thread
// (schedule(static)), which is how the reference's single-threaded solver would
// be deployed across host cores.
#include <omp.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <vector>

#include "kkt_oracle.hpp"
#include "theta_oracle.hpp"
#include "riccati_oracle.hpp"

using namespace sipoc_oracle;

extern "C" {

// Returns the LQR::FactorStatus-valued topology status; fills the CSR/order
// arrays (sizes E+2, E, E+1, E+1) when valid.
int oracle_compile_topology(int E, int root, const int *parents,
                            const int *children, int *child_offsets,
                            int *child_edges, int *preorder, int *postorder) {
  Tree t{E, root, parents, children};
  CompiledTree c = compile_tree(t);
  if (c.status != SUCCESS) return c.status;
  for (int i = 0; i < E + 2; ++i) child_offsets[i] = c.child_offsets[i];
  for (int i = 0; i < E; ++i) child_edges[i] = c.child_edges[i];
  for (int i = 0; i < E + 1; ++i) {
    preorder[i] = c.preorder[i];
    postorder[i] = c.postorder[i];
  }
  return SUCCESS;
}

// out[0..6] = per-problem element counts of the (n x n), (n), (n x m), (m x m),
// (m), A and B flat arrays.
void oracle_lqr_sizes(int E, int root, const int *parents, const int *children,
                      const int *state_dims, const int *control_dims,
                      int64_t *out) {
  Tree t{E, root, parents, children};
  FlatLayout L = make_layout(t, state_dims, control_dims);
  out[0] = L.nn_off[E + 1];
  out[1] = L.n_off[E + 1];
  out[2] = L.nm_off[E];
  out[3] = L.mm_off[E];
  out[4] = L.m_off[E];
  out[5] = L.a_off[E];
  out[6] = L.b_off[E];
}

// mode bit 0: factor, bit 1: solve (solve implies a factor in the same call).
// status[b] gets the FactorStatus of problem b; x/u/y are written only for
// problems whose factor succeeded.  residual (nullable) gets the KKT residual
// 2-norm per problem.  seconds (nullable) gets the wall time of the parallel
// factor(+solve) region, workspaces allocated outside it.
int oracle_lqr_batch(int E, int root, const int *parents, const int *children,
                     const int *state_dims, const int *control_dims,
                     int64_t batch, const double *Q, const double *M,
                     const double *R, const double *q, const double *r,
                     const double *A, const double *B, const double *c,
                     const double *delta, double *x, double *u, double *y,
                     int *status, double *residual, int mode, int repeats,
                     int nthreads, double *seconds) {
  Tree t{E, root, parents, children};
  CompiledTree ct = compile_tree(t);
  if (ct.status != SUCCESS) {
    for (int64_t b = 0; b < batch; ++b) status[b] = ct.status;
    return ct.status;
  }
  FlatLayout L = make_layout(t, state_dims, control_dims);
  const int64_t s_nn = L.nn_off[E + 1], s_n = L.n_off[E + 1], s_nm = L.nm_off[E],
                s_mm = L.mm_off[E], s_m = L.m_off[E], s_a = L.a_off[E],
                s_b = L.b_off[E];
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  std::vector<LqrWorkspace> ws(nthreads);
  for (auto &w : ws) w.reserve(L, ct);
  if (repeats < 1) repeats = 1;

  const auto t0 = std::chrono::steady_clock::now();
  for (int rep = 0; rep < repeats; ++rep) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t b = 0; b < batch; ++b) {
      LqrWorkspace &w = ws[omp_get_thread_num()];
      LqrInput in{Q + b * s_nn, M + b * s_nm, R + b * s_mm, q + b * s_n,
                  r + b * s_m,  A + b * s_a,  B + b * s_b,  c + b * s_n,
                  delta + b * s_n};
      const Status st = lqr_factor(ct, L, in, w);
      status[b] = st;
      if ((mode & 2) && st == SUCCESS) {
        LqrOutput out{x + b * s_n, u + b * s_m, y + b * s_n};
        lqr_solve(ct, L, in, w, out);
      }
    }
  }
  const auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();

  if (residual && (mode & 2)) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t b = 0; b < batch; ++b) {
      if (status[b] != SUCCESS) {
        residual[b] = -1.0;
        continue;
      }
      LqrInput in{Q + b * s_nn, M + b * s_nm, R + b * s_mm, q + b * s_n,
                  r + b * s_m,  A + b * s_a,  B + b * s_b,  c + b * s_n,
                  delta + b * s_n};
      LqrOutput out{x + b * s_n, u + b * s_m, y + b * s_n};
      residual[b] = lqr_residual_norm(ct, L, in, out);
    }
  }
  return SUCCESS;
}

// Residual only, for outputs computed elsewhere (the CUDA path).
int oracle_lqr_residual_batch(int E, int root, const int *parents,
                              const int *children, const int *state_dims,
                              const int *control_dims, int64_t batch,
                              const double *Q, const double *M, const double *R,
                              const double *q, const double *r, const double *A,
                              const double *B, const double *c,
                              const double *delta, const double *x,
                              const double *u, const double *y,
                              double *residual) {
  Tree t{E, root, parents, children};
  CompiledTree ct = compile_tree(t);
  if (ct.status != SUCCESS) return ct.status;
  FlatLayout L = make_layout(t, state_dims, control_dims);
  const int64_t s_nn = L.nn_off[E + 1], s_n = L.n_off[E + 1], s_nm = L.nm_off[E],
                s_mm = L.mm_off[E], s_m = L.m_off[E], s_a = L.a_off[E],
                s_b = L.b_off[E];
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < batch; ++b) {
    LqrInput in{Q + b * s_nn, M + b * s_nm, R + b * s_mm, q + b * s_n,
                r + b * s_m,  A + b * s_a,  B + b * s_b,  c + b * s_n,
                delta + b * s_n};
    LqrOutput out{const_cast<double *>(x + b * s_n), const_cast<double *>(u + b * s_m),
                  const_cast<double *>(y + b * s_n)};
    residual[b] = lqr_residual_norm(ct, L, in, out);
  }
  return SUCCESS;
}

// out[0..2] = x_dim, y_dim, z_dim; out[3..15] = per-problem element counts of
// node_hxx, node_jc, node_jg, edge_hxx, edge_hxu, edge_huu, edge_A, edge_B,
// edge_jcx, edge_jcu, edge_jgx, edge_jgu; out[15] unused.
void oracle_kkt_sizes(int E, int root, const int *parents, const int *children,
                      const int *state_dims, const int *control_dims,
                      const int *node_c, const int *node_g, const int *edge_c,
                      const int *edge_g, int64_t *out) {
  Tree t{E, root, parents, children};
  ConstraintDims cd{node_c, node_g, edge_c, edge_g};
  KktLayout K = make_kkt_layout(t, state_dims, control_dims, cd);
  int64_t hxx_edge = 0;
  for (int e = 0; e < E; ++e) {
    const int np = state_dims[parents[e]];
    hxx_edge += np * np;
  }
  out[0] = K.x_dim;
  out[1] = K.y_dim;
  out[2] = K.z_dim;
  out[3] = K.lqr.nn_off[E + 1];
  out[4] = K.jc_node_off[E + 1];
  out[5] = K.jg_node_off[E + 1];
  out[6] = hxx_edge;
  out[7] = K.lqr.nm_off[E];
  out[8] = K.lqr.mm_off[E];
  out[9] = K.lqr.a_off[E];
  out[10] = K.lqr.b_off[E];
  out[11] = K.jcx_off[E];
  out[12] = K.jcu_off[E];
  out[13] = K.jgx_off[E];
  out[14] = K.jgu_off[E];
  out[15] = 0;
}

// Offsets of the flat-vector wire format (types.cpp:24-64): arrays sized
// E+1 / E as appropriate.
void oracle_kkt_offsets(int E, int root, const int *parents, const int *children,
                        const int *state_dims, const int *control_dims,
                        const int *node_c, const int *node_g, const int *edge_c,
                        const int *edge_g, int *x_state, int *x_control,
                        int *y_dyn, int *y_node_c, int *y_edge_c, int *z_node,
                        int *z_edge) {
  Tree t{E, root, parents, children};
  ConstraintDims cd{node_c, node_g, edge_c, edge_g};
  KktLayout K = make_kkt_layout(t, state_dims, control_dims, cd);
  for (int i = 0; i <= E; ++i) {
    x_state[i] = K.x_state[i];
    y_dyn[i] = K.y_dyn[i];
    y_node_c[i] = K.y_node_c[i];
    z_node[i] = K.z_node[i];
  }
  for (int e = 0; e < E; ++e) {
    x_control[e] = K.x_control[e];
    y_edge_c[e] = K.y_edge_c[e];
    z_edge[e] = K.z_edge[e];
  }
}

// mode bit 0: factor, bit 1: solve, bit 2: residual norm ||K sol - b||_2 via
// kkt_apply (tests/variable_dimensions_test.cpp:159-180).  ok[b] = 1 when
// CallbackProvider::factor would return true.
int oracle_kkt_batch(int E, int root, const int *parents, const int *children,
                     const int *state_dims, const int *control_dims,
                     const int *node_c, const int *node_g, const int *edge_c,
                     const int *edge_g, int64_t batch, const double *node_hxx,
                     const double *node_jc, const double *node_jg,
                     const double *edge_hxx, const double *edge_hxu,
                     const double *edge_huu, const double *edge_A,
                     const double *edge_B, const double *edge_jcx,
                     const double *edge_jcu, const double *edge_jgx,
                     const double *edge_jgu, const double *w, const double *r1,
                     const double *r2, const double *r3, const double *b,
                     double *sol, int *ok, int *lqr_status, double *residual,
                     int mode, int repeats, int nthreads, double *seconds) {
  Tree t{E, root, parents, children};
  CompiledTree ct = compile_tree(t);
  if (ct.status != SUCCESS) {
    for (int64_t i = 0; i < batch; ++i) ok[i] = 0;
    return ct.status;
  }
  ConstraintDims cd{node_c, node_g, edge_c, edge_g};
  KktLayout K = make_kkt_layout(t, state_dims, control_dims, cd);
  int64_t sz[16];
  oracle_kkt_sizes(E, root, parents, children, state_dims, control_dims, node_c,
                   node_g, edge_c, edge_g, sz);
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  std::vector<KktWorkspace> ws(nthreads);
  for (auto &wk : ws) wk.reserve(K, ct);
  if (repeats < 1) repeats = 1;
  const int64_t kd = K.kkt_dim;

  const auto t0 = std::chrono::steady_clock::now();
  for (int rep = 0; rep < repeats; ++rep) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t i = 0; i < batch; ++i) {
      KktWorkspace &wk = ws[omp_get_thread_num()];
      KktModel mdl{node_hxx + i * sz[3],  node_jc + i * sz[4],
                   node_jg + i * sz[5],   edge_hxx + i * sz[6],
                   edge_hxu + i * sz[7],  edge_huu + i * sz[8],
                   edge_A + i * sz[9],    edge_B + i * sz[10],
                   edge_jcx + i * sz[11], edge_jcu + i * sz[12],
                   edge_jgx + i * sz[13], edge_jgu + i * sz[14]};
      int st = -1;
      const bool good = kkt_factor(ct, K, mdl, w + i * sz[2], r1 + i * sz[0],
                                   r2 + i * sz[1], r3 + i * sz[2], wk, &st);
      ok[i] = good ? 1 : 0;
      if (lqr_status) lqr_status[i] = st;
      if ((mode & 2) && good) {
        kkt_solve(ct, K, mdl, b + i * kd, sol + i * kd, wk);
      }
    }
  }
  const auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();

  if ((mode & 4) && residual) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t i = 0; i < batch; ++i) {
      if (!ok[i]) {
        residual[i] = -1.0;
        continue;
      }
      KktModel mdl{node_hxx + i * sz[3],  node_jc + i * sz[4],
                   node_jg + i * sz[5],   edge_hxx + i * sz[6],
                   edge_hxu + i * sz[7],  edge_huu + i * sz[8],
                   edge_A + i * sz[9],    edge_B + i * sz[10],
                   edge_jcx + i * sz[11], edge_jcu + i * sz[12],
                   edge_jgx + i * sz[13], edge_jgu + i * sz[14]};
      std::vector<double> prod(kd, 0.0);
      const double *s = sol + i * kd;
      kkt_apply(ct, K, mdl, w + i * sz[2], r1 + i * sz[0], r2 + i * sz[1],
                r3 + i * sz[2], s, s + K.x_dim, s + K.x_dim + K.y_dim,
                prod.data(), prod.data() + K.x_dim,
                prod.data() + K.x_dim + K.y_dim);
      double sq = 0.0;
      for (int64_t j = 0; j < kd; ++j) {
        const double d = prod[j] - b[i * kd + j];
        sq += d * d;
      }
      residual[i] = std::sqrt(sq);
    }
  }
  return SUCCESS;
}

// y += K x for one batch (add_Kx_to_y); x and y are [batch][kkt_dim].
int oracle_kkt_apply_batch(int E, int root, const int *parents,
                           const int *children, const int *state_dims,
                           const int *control_dims, const int *node_c,
                           const int *node_g, const int *edge_c,
                           const int *edge_g, int64_t batch,
                           const double *node_hxx, const double *node_jc,
                           const double *node_jg, const double *edge_hxx,
                           const double *edge_hxu, const double *edge_huu,
                           const double *edge_A, const double *edge_B,
                           const double *edge_jcx, const double *edge_jcu,
                           const double *edge_jgx, const double *edge_jgu,
                           const double *w, const double *r1, const double *r2,
                           const double *r3, const double *x, double *y) {
  Tree t{E, root, parents, children};
  CompiledTree ct = compile_tree(t);
  if (ct.status != SUCCESS) return ct.status;
  ConstraintDims cd{node_c, node_g, edge_c, edge_g};
  KktLayout K = make_kkt_layout(t, state_dims, control_dims, cd);
  int64_t sz[16];
  oracle_kkt_sizes(E, root, parents, children, state_dims, control_dims, node_c,
                   node_g, edge_c, edge_g, sz);
  const int64_t kd = K.kkt_dim;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < batch; ++i) {
    KktModel mdl{node_hxx + i * sz[3],  node_jc + i * sz[4],
                 node_jg + i * sz[5],   edge_hxx + i * sz[6],
                 edge_hxu + i * sz[7],  edge_huu + i * sz[8],
                 edge_A + i * sz[9],    edge_B + i * sz[10],
                 edge_jcx + i * sz[11], edge_jcu + i * sz[12],
                 edge_jgx + i * sz[13], edge_jgu + i * sz[14]};
    const double *xs = x + i * kd;
    double *ys = y + i * kd;
    kkt_apply(ct, K, mdl, w + i * sz[2], r1 + i * sz[0], r2 + i * sz[1],
              r3 + i * sz[2], xs, xs + K.x_dim, xs + K.x_dim + K.y_dim, ys,
              ys + K.x_dim, ys + K.x_dim + K.y_dim);
  }
  return SUCCESS;
}

// ---- theta (global / Schur variables): theta_oracle.hpp ------------------------------
// model: the 12 stagewise block arrays in KktModel order; theta: the 10 arrays in ThetaModel
// order; every array [batch][size].  r1 is [batch][x_dim + p]; b, sol, x, y are
// [batch][kkt_dim + p] in the full layout [x_s, theta | y | z].
// mode: 1 = factor, 2 = + solve, 8 = apply y += K x (b = x, sol = y) instead of factor / solve.
void oracle_kkt_theta_sizes(int E, int root, const int *parents, const int *children,
                            const int *state_dims, const int *control_dims, const int *node_c,
                            const int *node_g, const int *edge_c, const int *edge_g, int p,
                            int64_t *out) {
  Tree t{E, root, parents, children};
  ConstraintDims cd{node_c, node_g, edge_c, edge_g};
  KktLayout K = make_kkt_layout(t, state_dims, control_dims, cd);
  long long z[10];
  theta_sizes(K, t, p, z);
  for (int i = 0; i < 10; ++i) out[i] = z[i];
}

int oracle_kkt_theta_batch(int E, int root, const int *parents, const int *children,
                           const int *state_dims, const int *control_dims, const int *node_c,
                           const int *node_g, const int *edge_c, const int *edge_g, int p,
                           int64_t batch, const double *const *model,
                           const double *const *theta, const double *w, const double *r1,
                           const double *r2, const double *r3, const double *b, double *sol,
                           int *ok, int mode) {
  Tree t{E, root, parents, children};
  CompiledTree ct = compile_tree(t);
  if (ct.status != SUCCESS) {
    for (int64_t i = 0; i < batch && ok; ++i) ok[i] = 0;
    return ct.status;
  }
  ConstraintDims cd{node_c, node_g, edge_c, edge_g};
  KktLayout K = make_kkt_layout(t, state_dims, control_dims, cd);
  int64_t sz[16];
  oracle_kkt_sizes(E, root, parents, children, state_dims, control_dims, node_c, node_g, edge_c,
                   edge_g, sz);
  long long tz[10];
  theta_sizes(K, t, p, tz);
  const int64_t kd = K.kkt_dim + p, xd = K.x_dim + p;
  const int nthreads = omp_get_max_threads();
  std::vector<ThetaWorkspace> ws(nthreads);
  for (auto &wk : ws) wk.kkt.reserve(K, ct);
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t i = 0; i < batch; ++i) {
    ThetaWorkspace &wk = ws[omp_get_thread_num()];
    KktModel mdl{model[0] + i * sz[3],   model[1] + i * sz[4],   model[2] + i * sz[5],
                 model[3] + i * sz[6],   model[4] + i * sz[7],   model[5] + i * sz[8],
                 model[6] + i * sz[9],   model[7] + i * sz[10],  model[8] + i * sz[11],
                 model[9] + i * sz[12],  model[10] + i * sz[13], model[11] + i * sz[14]};
    ThetaModel tm{theta[0] + i * tz[0], theta[1] + i * tz[1], theta[2] + i * tz[2],
                  theta[3] + i * tz[3], theta[4] + i * tz[4], theta[5] + i * tz[5],
                  theta[6] + i * tz[6], theta[7] + i * tz[7], theta[8] + i * tz[8],
                  theta[9] + i * tz[9]};
    const double *wi = w + i * sz[2], *r1i = r1 + i * xd, *r2i = r2 + i * sz[1],
                 *r3i = r3 + i * sz[2];
    if (mode & 8) {
      theta_apply(ct, t, K, p, mdl, tm, wi, r1i, r2i, r3i, b + i * kd, sol + i * kd);
      continue;
    }
    const bool good = theta_factor(ct, t, K, p, mdl, tm, wi, r1i, r2i, r3i, wk);
    ok[i] = good ? 1 : 0;
    if ((mode & 2) && good) theta_solve(ct, K, p, mdl, b + i * kd, sol + i * kd, wk);
  }
  return SUCCESS;
}

// ---- model-callback scatter: sip_optimal_control.cpp:44-123 --------------------------
// The node / edge values of one model evaluation scattered into the flat objective, gradient,
// equality and inequality vectors the interior-point loop reads (get_f / get_grad_f / get_c /
// get_g).  vals: twelve arrays, problem-major [batch][size]:
//   0 node f [N]            1 node df_dx [sum n]        2 node df_dtheta [N p]
//   3 node c [sum node_c]   4 node g [sum node_g]
//   5 edge f [E]            6 edge df_dx [sum n_parent] 7 edge df_du [sum m]
//   8 edge df_dtheta [E p]  9 edge dyn_res [sum n_child] 10 edge c [sum edge_c]
//   11 edge g [sum edge_g]
// x [batch][x_dim + p], x0 [batch][n_root]; f [batch], grad [batch][x_dim + p],
// c [batch][y_dim], g [batch][z_dim].  new_x == 0 computes f only (:54).
int oracle_model_scatter_batch(int E, int root, const int *parents, const int *children,
                               const int *state_dims, const int *control_dims,
                               const int *node_c, const int *node_g, const int *edge_c,
                               const int *edge_g, int p, int64_t batch,
                               const double *const *vals, const double *x, const double *x0,
                               int new_x, double *f, double *grad, double *c, double *g) {
  Tree t{E, root, parents, children};
  ConstraintDims cd{node_c, node_g, edge_c, edge_g};
  KktLayout K = make_kkt_layout(t, state_dims, control_dims, cd);
  const int N = E + 1;
  std::vector<int> n_off(N + 1, 0), nc_off(N + 1, 0), ng_off(N + 1, 0);
  std::vector<int> pn_off(E + 1, 0), cn_off(E + 1, 0), m_off(E + 1, 0), ec_off(E + 1, 0),
      eg_off(E + 1, 0);
  for (int i = 0; i < N; ++i) {
    n_off[i + 1] = n_off[i] + state_dims[i];
    nc_off[i + 1] = nc_off[i] + node_c[i];
    ng_off[i + 1] = ng_off[i] + node_g[i];
  }
  for (int e = 0; e < E; ++e) {
    pn_off[e + 1] = pn_off[e] + state_dims[parents[e]];
    cn_off[e + 1] = cn_off[e] + state_dims[children[e]];
    m_off[e + 1] = m_off[e] + control_dims[e];
    ec_off[e + 1] = ec_off[e] + edge_c[e];
    eg_off[e + 1] = eg_off[e] + edge_g[e];
  }
  const int64_t size[12] = {N,         n_off[N],  int64_t(N) * p, nc_off[N], ng_off[N], E,
                            pn_off[E], m_off[E],  int64_t(E) * p, cn_off[E], ec_off[E], eg_off[E]};
  const int xd = K.x_dim + p;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < batch; ++i) {
    const double *v[12];
    for (int k = 0; k < 12; ++k) v[k] = vals[k] + i * size[k];
    const double *xi = x + i * xd;
    double *gi = grad + i * xd, *ci = c + i * K.y_dim, *zi = g + i * K.z_dim;
    double fi = 0.0;                                              // :44-50
    for (int node = 0; node < N; ++node) fi += v[0][node];
    for (int e = 0; e < E; ++e) fi += v[5][e];
    f[i] = fi;
    if (!new_x) continue;                                         // :52
    for (int k = 0; k < xd; ++k) gi[k] = 0.0;                     // :54
    for (int node = 0; node < N; ++node) {                        // :55-66
      for (int row = 0; row < state_dims[node]; ++row)
        gi[K.x_state[node] + row] += v[1][n_off[node] + row];
      for (int row = 0; row < p; ++row) gi[K.x_dim + row] += v[2][node * p + row];
    }
    for (int e = 0; e < E; ++e) {                                 // :67-84
      const int parent = parents[e];
      for (int row = 0; row < state_dims[parent]; ++row)
        gi[K.x_state[parent] + row] += v[6][pn_off[e] + row];
      for (int row = 0; row < control_dims[e]; ++row)
        gi[K.x_control[e] + row] += v[7][m_off[e] + row];
      for (int row = 0; row < p; ++row) gi[K.x_dim + row] += v[8][e * p + row];
    }
    for (int row = 0; row < state_dims[root]; ++row)              // :88-94
      ci[K.y_dyn[root] + row] = x0[i * state_dims[root] + row] - xi[K.x_state[root] + row];
    for (int node = 0; node < N; ++node)                          // :95-99
      for (int r = 0; r < node_c[node]; ++r) ci[K.y_node_c[node] + r] = v[3][nc_off[node] + r];
    for (int e = 0; e < E; ++e) {                                 // :100-108
      const int child = children[e];
      for (int r = 0; r < state_dims[child]; ++r) ci[K.y_dyn[child] + r] = v[9][cn_off[e] + r];
      for (int r = 0; r < edge_c[e]; ++r) ci[K.y_edge_c[e] + r] = v[10][ec_off[e] + r];
    }
    for (int node = 0; node < N; ++node)                          // :112-116
      for (int r = 0; r < node_g[node]; ++r) zi[K.z_node[node] + r] = v[4][ng_off[node] + r];
    for (int e = 0; e < E; ++e)                                   // :117-121
      for (int r = 0; r < edge_g[e]; ++r) zi[K.z_edge[e] + r] = v[11][eg_off[e] + r];
  }
  return SUCCESS;
}

int oracle_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
