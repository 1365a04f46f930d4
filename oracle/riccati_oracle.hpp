// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Eigen-free CPU restatement of the reference's regularized tree-LQR
// (joaospinto/sip_optimal_control, sip_optimal_control/lqr.cpp) used as the
// parity checker for the CUDA path and as the "port" CPU baseline.  Nothing in
// the product (sip_optimal_control_b200/, include/) may link or call this;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs do.
//
// Pinning: the reference itself cannot be compiled in this environment (it
// needs Eigen 3.4.0, absent here and unreachable offline), so the restatement
// is pinned on the closed-form fixtures of the reference's own tests
// (tests/lqr_test.cpp, tests/variable_dimensions_test.cpp) restated verbatim
// under tests/, with the reference's own acceptance bars (KKT residual < 1e-12,
// dense-KKT agreement 1e-10, exact status codes and traversal orders).
//
// Third-party arithmetic restated here: Eigen 3.4.0 (MODULE.bazel:16) —
// unblocked lower LLT (Eigen/src/Cholesky/LLT.h llt_inplace<Lower>::unblocked,
// failure when a pivot is <= 0), triangular solveInPlace (column-wise
// forward / back substitution) and dense products (plain triple loops; the
// summation order inside Eigen's GEBP kernels is not reproduced, which moves
// results by O(eps)).
#pragma once

#include <vector>

namespace sipoc_oracle {

// Same numeric values as LQR::FactorStatus (lqr.hpp:68-74).
enum Status : int {
  SUCCESS = 0,
  INVALID_DELTA = 1,
  F_FACTORIZATION_FAILURE = 2,
  G_FACTORIZATION_FAILURE = 3,
  INVALID_TOPOLOGY = 4,
};

// Rooted tree, one edge per non-root node (lqr.hpp:5-22).
struct Tree {
  int num_edges = 0;
  int root = 0;
  const int *parents = nullptr;
  const int *children = nullptr;
  int num_nodes() const { return num_edges + 1; }
};

// CSR children + DFS orders (lqr.cpp:563-631).
struct CompiledTree {
  Status status = INVALID_TOPOLOGY;
  std::vector<int> child_offsets;  // num_nodes + 1
  std::vector<int> child_edges;    // num_edges
  std::vector<int> parents;        // num_edges
  std::vector<int> children;       // num_edges
  std::vector<int> preorder;       // num_nodes
  std::vector<int> postorder;      // num_nodes
};

CompiledTree compile_tree(const Tree &tree);

// Offsets of every node / edge block inside the per-problem flat arrays.
// Flat arrays concatenate the column-major blocks in node (resp. edge) index
// order; this is the "flat per-problem index" shared with the CUDA engine,
// whose HBM layout is the transpose [flat index][problem].
struct FlatLayout {
  int num_edges = 0;
  std::vector<int> n;      // state dim per node
  std::vector<int> m;      // control dim per edge
  std::vector<int> nn_off; // Q / V / F blocks (n_i x n_i), per node, +total
  std::vector<int> n_off;  // q, c, delta, x, y (n_i), per node, +total
  std::vector<int> nm_off; // M (n_parent x m_e), per edge, +total
  std::vector<int> mm_off; // R (m_e x m_e), per edge, +total
  std::vector<int> m_off;  // r, u (m_e), per edge, +total
  std::vector<int> a_off;  // A (n_child x n_parent), per edge, +total
  std::vector<int> b_off;  // B (n_child x m_e), per edge, +total
};

FlatLayout make_layout(const Tree &tree, const int *state_dims,
                       const int *control_dims);

// One problem, flat arrays (see FlatLayout).
struct LqrInput {
  const double *Q, *M, *R, *q, *r, *A, *B, *c, *delta;
};

struct LqrOutput {
  double *x, *u, *y;
};

// Everything LQR::Workspace keeps between factor and solve (lqr.hpp:109-119).
struct LqrWorkspace {
  std::vector<double> W, K, V, G_factor, F_factor, sqrt_delta, sqrt_delta_inv,
      k, v;
  std::vector<int> w_off, k_off;  // per edge: W (n_c x n_c), K (m x n_p)
  // single-edge scratch (lqr.hpp:121-127)
  std::vector<double> H, F, f, g, h;
  void reserve(const FlatLayout &layout, const CompiledTree &tree);
};

// lqr.cpp:645-731
Status lqr_factor(const CompiledTree &tree, const FlatLayout &layout,
                  const LqrInput &in, LqrWorkspace &ws);

// lqr.cpp:735-871
void lqr_solve(const CompiledTree &tree, const FlatLayout &layout,
               const LqrInput &in, LqrWorkspace &ws, const LqrOutput &out);

// Residual of the KKT system the LQR solves (tests/lqr_test.cpp:152-186 for
// chains, :371-409 for trees); returns the 2-norm.
double lqr_residual_norm(const CompiledTree &tree, const FlatLayout &layout,
                         const LqrInput &in, const LqrOutput &out);

// --- small dense kernels, exposed for unit tests -------------------------

// In-place unblocked lower Cholesky, column-major, leading dimension n.
// Returns false when a pivot is <= 0 (Eigen LLT NumericalIssue).
bool cholesky_lower_inplace(double *a, int n);
// Solve L X = B then L^T X = B in place; B is n x nrhs column-major.
void cholesky_solve_inplace(const double *l, int n, double *b, int nrhs);

}  // namespace sipoc_oracle
