// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Eigen-free CPU restatement of the theta_dim == 0 parts of the reference's
// Newton-KKT -> regularized-LQR reduction (sip_optimal_control/helpers.cpp:
// CallbackProvider::factor :242-370, single-RHS solve_stagewise_kkt_matrix
// :749-894, add_Kx_to_y :953-977 with the theta == 0 branches of :979-1368)
// and of the flat-vector offset tables (types.cpp:24-64).  Same role and same
// pinning as riccati_oracle.hpp: it is the checker, never the product.
#pragma once

#include <vector>

#include "riccati_oracle.hpp"

namespace sipoc_oracle {

// Per-node / per-edge constraint dimensions (lqr.hpp:24-33); null = all zero.
struct ConstraintDims {
  const int *node_c = nullptr;
  const int *node_g = nullptr;
  const int *edge_c = nullptr;
  const int *edge_g = nullptr;
};

// Offsets into the flat x / y / z vectors (types.cpp:24-64) and into the
// per-problem flat model-block arrays.
struct KktLayout {
  FlatLayout lqr;
  std::vector<int> node_c, node_g, edge_c, edge_g;
  int x_dim = 0, y_dim = 0, z_dim = 0, kkt_dim = 0;
  std::vector<int> x_state, x_control;          // per node / per edge
  std::vector<int> y_dyn, y_node_c, y_edge_c;   // per node / node / edge
  std::vector<int> z_node, z_edge;              // per node / per edge
  // model blocks (column-major), per node then total
  std::vector<int> jc_node_off;  // dc_dx  (c_i x n_i)
  std::vector<int> jg_node_off;  // dg_dx  (g_i x n_i)
  // per edge then total
  std::vector<int> jcx_off;      // dc_dx  (c_e x n_p)
  std::vector<int> jcu_off;      // dc_du  (c_e x m)
  std::vector<int> jgx_off;      // dg_dx  (g_e x n_p)
  std::vector<int> jgu_off;      // dg_du  (g_e x m)
};

KktLayout make_kkt_layout(const Tree &tree, const int *state_dims,
                          const int *control_dims, const ConstraintDims &cd);

// The ModelCallbackOutput blocks the reduction reads (types.hpp:48-89), flat.
// Node blocks: d2L_dx2 uses FlatLayout::nn_off.  Edge blocks: d2L_dx2 is
// (n_p x n_p) at hxx_edge_off, d2L_dxdu at lqr.nm_off, d2L_du2 at lqr.mm_off,
// ddyn_dx at lqr.a_off, ddyn_du at lqr.b_off.
struct KktModel {
  const double *node_hxx, *node_jc, *node_jg;
  const double *edge_hxx, *edge_hxu, *edge_huu, *edge_A, *edge_B;
  const double *edge_jcx, *edge_jcu, *edge_jgx, *edge_jgu;
};

struct KktWorkspace {
  std::vector<int> hxx_edge_off;  // per edge (n_p x n_p), +total
  std::vector<double> Q_mod, M_mod, R_mod, q_mod, r_mod, c_mod, dyn_r2;
  std::vector<double> node_c_r2_inv, edge_c_r2_inv, node_mod_w_inv, edge_mod_w_inv;
  std::vector<int> node_c_off, node_g_off, edge_c_off, edge_g_off;
  std::vector<double> x, u, y;  // LQR outputs before scatter
  LqrWorkspace lqr;
  void reserve(const KktLayout &layout, const CompiledTree &tree);
};

// helpers.cpp:242-370; returns false exactly where the reference does
// (non-positive r2 / w + r3, LQR factor failure).  lqr_status (optional)
// receives the LQR status when the reduction itself succeeded.
bool kkt_factor(const CompiledTree &tree, const KktLayout &layout,
                const KktModel &model, const double *w, const double *r1,
                const double *r2, const double *r3, KktWorkspace &ws,
                int *lqr_status);

// helpers.cpp:749-894 (single right-hand side).
void kkt_solve(const CompiledTree &tree, const KktLayout &layout,
               const KktModel &model, const double *b, double *sol,
               KktWorkspace &ws);

// helpers.cpp:953-977 (theta == 0): y += K(w, r1, r2, r3) x, block-wise.
void kkt_apply(const CompiledTree &tree, const KktLayout &layout,
               const KktModel &model, const double *w, const double *r1,
               const double *r2, const double *r3, const double *x_x,
               const double *x_y, const double *x_z, double *y_x, double *y_y,
               double *y_z);

}  // namespace sipoc_oracle
