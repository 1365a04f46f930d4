// Synthetic workload generator (bench / test tooling, not part of the solve).
//
// Reproduces the DISTRIBUTION of the reference's LQRProblem generator
// (benchmarks/lqr_benchmark.cpp:61-96, entries drawn column-major :119-129):
//   A = I + 0.05 N(0,1)   B = 0.1 N(0,1)   M = 0
//   R = Z'Z + (1 + 1e-2) I      Q = Z'Z + 1e-3 I (terminal node included)
//   q, r, c ~ N(0,1)            delta = 1e-3 + 1e-1 U(0,1)
// The reference draws from std::mt19937 through std::normal_distribution,
// whose stream is implementation-defined, so only the distribution is
// normative.  Here every scalar is a pure function of
// (seed, array id, global problem index, flat element index) through a
// splitmix64-style counter hash, so any shard of the batch can be generated
// independently on its own GPU.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace sipoc {

// Returns the number of kernels launched.
int launch_generate_lqr_benchmark(uint64_t seed, int64_t problem_offset, int num_edges, int n,
                                  int m, int64_t batch, int64_t ld, double *Q, double *M,
                                  double *R, double *q, double *r, double *A, double *B,
                                  double *c, double *delta, cudaStream_t stream);

}  // namespace sipoc
