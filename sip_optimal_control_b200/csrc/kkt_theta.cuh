// Global ("theta") variables of the Newton-KKT system: the Schur-complement layer the
// reference builds on top of the stagewise solve (helpers.cpp:190-240 form_theta_jacobian,
// :372-407 Schur factor, :902-951 solve, theta branches of :1019-1368 operator).
//
//   K = [ K_s   J  ]   K_s = the stagewise KKT matrix the LQR path factors,
//       [ J'    H_tt + R1_tt ]   J = d(stagewise rows)/d(theta), one column per theta entry.
//
// factor:  K_s (existing path) ; J ; Kinv J = p stagewise solves against the kept
//          factorization (the reference's multi-RHS solve_stagewise_kkt_matrix, :422-747,
//          does the same statements on p columns at once) ; S = sum H_tt + R1_tt - J' Kinv J ;
//          Cholesky of S (pivot <= 0 -> factor returns false).
// solve:   sol_s = K_s^-1 b_s ; t = b_t - J' sol_s ; theta = S^-1 t ; sol_s -= (Kinv J) theta.
//
// Everything is kept in the FULL vector layout [x_0, u_0, ..., x_E, theta | y | z]: the
// stagewise kernels address y and z through DevTables::x_dim (which counts theta) and never
// touch the theta rows, so J, Kinv J and the solution need no compaction.
#pragma once

#include <cuda_runtime.h>

#include "structure.hpp"

namespace sipoc {

// theta blocks of ModelCallbackOutput (types.hpp:48-89), engine layout, column-major per
// block with p = theta_dim columns:
//   node  d2L_dxdtheta [n x p], dc_dtheta [c x p], dg_dtheta [g x p], d2L_dtheta2 [p x p]
//   edge  d2L_dxdtheta [n_parent x p], d2L_dudtheta [m x p], ddyn_dtheta [n_child x p],
//         dc_dtheta [c x p], dg_dtheta [g x p], d2L_dtheta2 [p x p]
struct KktThetaModel {
  const double *node_hxt, *node_jct, *node_jgt, *node_htt;
  const double *edge_hxt, *edge_hut, *edge_dynt, *edge_jct, *edge_jgt, *edge_htt;
};

// J [kkt_dim x p], written whole (zero where no block lands).
void launch_theta_jacobian(const DevTables &t, const KktThetaModel &m, double *J, int64_t batch,
                           int64_t ld, cudaStream_t s);
// S = sum H_tt + diag(r1 theta rows) - J' KinvJ (lower), then its Cholesky factor in place;
// ok[b] is cleared where a pivot is <= 0 (Eigen::LLT's failure rule).
void launch_theta_schur(const DevTables &t, const KktThetaModel &m, const double *r1,
                        const double *J, const double *KinvJ, double *S, int *ok, int64_t batch,
                        int64_t ld, cudaStream_t s);
// Given sol = K_s^-1 b on the stagewise rows: theta rows and the correction of the others.
// `tvec` is p doubles of scratch per problem.
void launch_theta_solve(const DevTables &t, const double *b, const double *J, const double *KinvJ,
                        const double *S, double *tvec, double *sol, int64_t batch, int64_t ld,
                        cudaStream_t s);
// theta terms of y += K x; `parts` as for launch_kkt_apply_parts (H, C, CT, G, GT, Reg bits).
void launch_theta_apply(const DevTables &t, const KktThetaModel &m, unsigned parts, const double *r1,
                        const double *in_x, const double *in_y, const double *in_z, double *out_x,
                        double *out_y, double *out_z, int64_t batch, int64_t ld, cudaStream_t s);

}  // namespace sipoc
