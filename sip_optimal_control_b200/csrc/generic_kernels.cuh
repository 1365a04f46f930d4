// Generic (any tree, any per-node / per-edge dimensions) CUDA kernels.
//
// Mapping: ONE THREAD PER PROBLEM.  Every per-problem array lives in HBM in the
// engine layout X[flat index * ld + b] (batch innermost), so the 32 threads of
// a warp — 32 neighbouring problems executing the same instruction on the same
// flat index — always touch 32 consecutive doubles (one 256-byte, fully
// coalesced request).  Control flow depends only on the shared structure
// tables, so warps never diverge.  Operation order mirrors the reference
// statement by statement (lqr.cpp:475-549, 645-871; helpers.cpp:242-370,
// 749-894, 953-1368), which makes this path the numerically strict one; the
// shape-specialised kernels in riccati_fast.cuh trade that for speed on
// uniform chains.
#pragma once

#include <cuda_runtime.h>

#include "structure.hpp"

namespace sipoc {

// Batch-innermost view of one problem's slice of an array.
struct GVec {
  double *p;
  size_t ld;
  __device__ __forceinline__ double &operator()(int off) const {
    return p[static_cast<size_t>(off) * ld];
  }
};
struct GCVec {
  const double *p;
  size_t ld;
  __device__ __forceinline__ double operator()(int off) const {
    return __ldg(p + static_cast<size_t>(off) * ld);
  }
};

struct LqrIn {
  const double *Q, *M, *R, *q, *r, *A, *B, *c, *delta;
};
struct LqrOut {
  double *x, *u, *y;
};
// Factorization kept between factor and solve (reference LQR::Workspace,
// lqr.hpp:109-127), engine layout.
struct LqrWs {
  double *W, *K, *V, *Gf, *Ff, *sd, *sdi, *k, *v;  // persistent
  double *H, *F, *f, *g, *h;                       // per-problem scratch
};

struct KktModel {
  const double *node_hxx, *node_jc, *node_jg;
  const double *edge_hxx, *edge_hxu, *edge_huu, *edge_A, *edge_B;
  const double *edge_jcx, *edge_jcu, *edge_jgx, *edge_jgu;
};
// Outputs of the reduction (reference Workspace::RegularizedLQRData,
// types.hpp:163-174), engine layout.
struct KktWs {
  double *Q_mod, *M_mod, *R_mod, *q_mod, *r_mod, *c_mod, *dyn_r2;
  double *node_c_r2_inv, *edge_c_r2_inv, *node_mod_w_inv, *edge_mod_w_inv;
  double *x, *u, *y;  // LQR outputs before the scatter into sol
};

void launch_generic_lqr_factor(const DevTables &t, const LqrIn &in, const LqrWs &ws,
                               int *status, int64_t batch, int64_t ld,
                               cudaStream_t stream);
void launch_generic_lqr_solve(const DevTables &t, const LqrIn &in, const LqrWs &ws,
                              const LqrOut &out, int64_t batch, int64_t ld,
                              cudaStream_t stream);
void launch_lqr_residual(const DevTables &t, const LqrIn &in, const LqrOut &out,
                         const int *status, double *residual_norm, double *stats,
                         int64_t batch, int64_t ld, cudaStream_t stream);

// Newton-KKT: reduction (factor prologue), rhs build, dual recovery, operator.
void launch_kkt_reduce(const DevTables &t, const KktModel &m, const double *w,
                       const double *r1, const double *r2, const double *r3,
                       const KktWs &ws, int *ok, int64_t batch, int64_t ld,
                       cudaStream_t stream);
void launch_kkt_finish_factor(const int *lqr_status, int *ok, int64_t batch,
                              cudaStream_t stream);
void launch_kkt_build_rhs(const DevTables &t, const KktModel &m, const KktWs &ws,
                          const double *b, int64_t batch, int64_t ld,
                          cudaStream_t stream);
void launch_kkt_recover(const DevTables &t, const KktModel &m, const KktWs &ws,
                        const double *b, double *sol, int64_t batch, int64_t ld,
                        cudaStream_t stream);
// Blocks of the KKT operator K = [H + R1, C', G'; C, -R2, 0; G, 0, -(W + R3)]
// (helpers.cpp:953-1368): add_Kx_to_y applies all of them, add_{H,C,CT,G,GT}x_to_y one each.
constexpr unsigned kKktH = 1u, kKktC = 2u, kKktCT = 4u, kKktG = 8u, kKktGT = 16u, kKktReg = 32u,
                   kKktAll = 63u;
void launch_kkt_apply_parts(const DevTables &t, const KktModel &m, unsigned parts,
                            const double *in_x, const double *in_y, const double *in_z,
                            double *out_x, double *out_y, double *out_z, int64_t batch,
                            int64_t ld, cudaStream_t s);
void launch_kkt_apply(const DevTables &t, const KktModel &m, const double *w,
                      const double *r1, const double *r2, const double *r3,
                      const double *x, double *y, int64_t batch, int64_t ld,
                      cudaStream_t stream);
void launch_kkt_residual(const DevTables &t, const double *Ksol, const double *b,
                         const int *ok, double *residual_norm, double *stats,
                         int64_t batch, int64_t ld, cudaStream_t stream);

// Layout conversion: problem-major [batch][size] <-> engine [size][ld].
// Variable-dimension chain -> uniform (np, mp) chain with decoupled padding, and the
// outputs back.  mask bit i selects array i of {Q, M, R, q, r, A, B, c, delta}.
void launch_pad_chain(const DevTables &t, const LqrIn &src, const LqrIn &dst, int np, int mp,
                      unsigned mask, int64_t batch, int64_t ld, cudaStream_t s);
void launch_unpad_chain(const DevTables &t, const LqrOut &src, const LqrOut &dst, int np, int mp,
                        int64_t batch, int64_t ld, cudaStream_t s);
void launch_pack(const double *src, double *dst, int64_t size, int64_t batch,
                 int64_t ld, cudaStream_t stream);
void launch_unpack(const double *src, double *dst, int64_t size, int64_t batch,
                   int64_t ld, cudaStream_t stream);
void launch_status_stats(const int *status, double *stats, int64_t batch,
                         cudaStream_t stream);
void launch_fill_int(int *dst, int value, int64_t count, cudaStream_t stream);

// Number of kernel launches each launcher above performs (for launch_count).
}  // namespace sipoc
