// Shape-specialised Newton-KKT -> LQR reduction for uniform chains (kkt_fast.cu).
#pragma once

#include <cuda_runtime.h>

#include "generic_kernels.cuh"

namespace sipoc {

// Same contract as launch_kkt_reduce (fills ok with 1, clears it where a
// regularization term is not positive, writes Q_mod / M_mod / R_mod / dyn_r2 and the
// four weight arrays).  max_rows = the largest node_c + node_g + edge_c + edge_g of
// any node and its child edge.  Launches 2 kernels.
using KktReduceFn = void (*)(const DevTables &, const KktModel &, const double *w,
                             const double *r1, const double *r2, const double *r3,
                             const KktWs &, int *ok, int64_t batch, int64_t ld, int max_rows,
                             cudaStream_t);

// nullptr when (n, m) is not instantiated.
KktReduceFn select_kkt_reduce(int n, int m);
// Dynamic shared memory of the reduction kernel: weights + Jacobian rows of 32 problems.
inline size_t kkt_reduce_smem_bytes(int n, int m, int max_rows) {
  return static_cast<size_t>(max_rows) * (n + m + 1) * 32 * sizeof(double);
}

// Same contract as launch_kkt_apply: y += K x on [x | y | z] vectors in the engine layout.
using KktApplyFn = void (*)(const DevTables &, const KktModel &, const double *w,
                            const double *r1, const double *r2, const double *r3,
                            const double *x, double *y, int64_t batch, int64_t ld, cudaStream_t);
KktApplyFn select_kkt_apply(int n, int m);

}  // namespace sipoc
