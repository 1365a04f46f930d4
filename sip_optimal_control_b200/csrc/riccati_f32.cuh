// Optional FP32 mode: factor + solve on uniform chains with n = 4 in single precision
// (riccati_f32.cu).  The same kernels instantiated on double serve as their own numerical
// control (tests hold that instantiation to the FP64 tolerance).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/sipoc.h"
#include "profile.hpp"

namespace sipoc {

template <class T>
struct LqrInT {
  const T *Q, *M, *R, *q, *r, *A, *B, *c, *delta;
};
template <class T>
struct LqrOutT {
  T *x, *u, *y;
};

bool f32_supports(int n, int m);
bool f32_fits_index(int n, int m, int T, int64_t ld);
// Elements per problem of the kept factorization + affine spill (P, K, v, k).
int64_t f32_store_elems(int n, int m, int T);
// Launch count, or -1 for an unsupported shape.  Arrays in the engine layout [flat][ld].
int launch_lqr_factor_solve_f32(int n, int m, const LqrInT<float> &in, const LqrOutT<float> &out,
                                int *status, float *store, int64_t batch, int64_t ld, int T,
                                Profiler *prof, cudaStream_t s);
int launch_lqr_factor_solve_thread_f64(int n, int m, const LqrInT<double> &in,
                                       const LqrOutT<double> &out, int *status, double *store,
                                       int64_t batch, int64_t ld, int T, Profiler *prof,
                                       cudaStream_t s);

}  // namespace sipoc
