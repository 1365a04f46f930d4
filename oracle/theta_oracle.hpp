// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Eigen-free CPU restatement of the theta (global / Schur variable) layer of the
// reference's CallbackProvider (sip_optimal_control/helpers.cpp): form_theta_jacobian
// :190-240, the Schur complement and its Cholesky factor in factor :372-407, solve :896-951,
// and the theta branches of add_{H,C,CT,G,GT}x_to_y :1019-1368.  The stagewise solves run
// through kkt_oracle's single-right-hand-side restatement, one column at a time: the
// reference's multi-RHS solve_stagewise_kkt_matrix (:422-747) performs the same statements
// on p columns at once.  Parity is pinned like kkt_oracle's (the reference's own fixture,
// tests/variable_dimensions_test.cpp:338-363: ||K sol - rhs|| < 1e-9, and a dense solve).
#pragma once

#include <vector>

#include "kkt_oracle.hpp"

namespace sipoc_oracle {

// theta blocks of ModelCallbackOutput (types.hpp:48-89), flat per problem, column-major
// with p = theta_dim columns; node blocks in node order, edge blocks in edge order.
struct ThetaModel {
  const double *node_hxt, *node_jct, *node_jgt, *node_htt;
  const double *edge_hxt, *edge_hut, *edge_dynt, *edge_jct, *edge_jgt, *edge_htt;
};

// Element counts of those ten arrays for one problem.
void theta_sizes(const KktLayout &K, const Tree &tree, int p, long long out[10]);

struct ThetaWorkspace {
  std::vector<double> J, KinvJ;        // [stagewise_kkt_dim x p], column-major
  std::vector<double> S, L;            // [p x p]
  std::vector<double> rhs, sol, t;     // stagewise rhs / solution, theta rhs
  KktWorkspace kkt;
};

// CallbackProvider::factor with theta_dim = p > 0.  r1 is indexed like the full x.
bool theta_factor(const CompiledTree &ct, const Tree &tree, const KktLayout &K, int p,
                  const KktModel &m, const ThetaModel &tm, const double *w, const double *r1,
                  const double *r2, const double *r3, ThetaWorkspace &ws);
// CallbackProvider::solve: b, sol are full vectors [x_s, theta | y | z].
void theta_solve(const CompiledTree &ct, const KktLayout &K, int p, const KktModel &m,
                 const double *b, double *sol, ThetaWorkspace &ws);
// add_Kx_to_y on full vectors.
void theta_apply(const CompiledTree &ct, const Tree &tree, const KktLayout &K, int p,
                 const KktModel &m, const ThetaModel &tm, const double *w, const double *r1,
                 const double *r2, const double *r3, const double *x, double *y);

}  // namespace sipoc_oracle
