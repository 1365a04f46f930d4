// Shape-specialised Riccati kernels for uniform chains (interface).
//
// A FastPlan is a set of kernels compiled for one (state_dim, control_dim)
// pair.  riccati_fast.cu instantiates n in {4, 6, 8} x m in {1, 2, 3, 4} (the
// reference benchmark grid below n = 16, lqr_benchmark.cpp:537-545) and the
// quadrotor shape (12, 4); riccati_cta.cu (16, 4), (32, 8) and (64, 24); every
// other uniform chain is padded up to the next of these with decoupled states /
// controls (api.cu, sipoc_create), small chains / trees with varying dims run the
// reference-order plans of riccati_strict.cu, and the generic thread-per-problem
// kernels take whatever is left.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "generic_kernels.cuh"
#include "profile.hpp"

namespace sipoc {

struct FastArgs {
  LqrIn in;
  // Problem-major copies [problem][flat] of the inputs (plans with
  // problem_major_inputs; pointers are null otherwise or for arrays not refreshed).
  LqrIn pm;
  LqrOut out;
  int *status;      // device int[ld] or nullptr
  double *store;    // factorization kept for later solves (W, K, LG per stage)
  double *scratch;  // per-problem spill of the fused / solve rollout (v, k)
  int64_t batch, ld;
  int num_edges;
  Profiler *prof;   // per-kernel event timing (may be disabled)
  const DevTables *tables = nullptr;  // topology, for the plans that serve trees
};

// Newton-KKT solve against the plan's kept factorization (uniform chains).
struct FastKktArgs {
  const DevTables *tables;
  const KktModel *model;
  const KktWs *ws;   // weights and dyn_r2 written by the reduction
  const double *b;   // rhs,      [x | y | z], engine layout
  double *sol;       // solution, [x | y | z], engine layout
  const double *store;
  double *scratch;
  int64_t batch, ld;
  int num_edges;
  Profiler *prof;
};

struct FastPlan {
  const char *name;
  int n, m;
  // doubles per problem the engine must provide
  int64_t (*store_elems)(int num_edges);
  int64_t (*scratch_elems)(int num_edges);
  // Each returns the number of kernels it launched.
  int (*factor)(const FastArgs &, cudaStream_t);
  int (*solve)(const FastArgs &, cudaStream_t);
  int (*factor_solve)(const FastArgs &, cudaStream_t);
  // Fused rhs build + affine sweep, rollout + dual recovery; nullptr = not provided.
  int (*kkt_solve)(const FastKktArgs &, cudaStream_t);
  // The backward kernel wants FastArgs::pm: one problem per CTA reads whole sectors
  // from a problem-major copy instead of 8 bytes per 32-byte sector of the
  // batch-interleaved layout.
  bool problem_major_inputs;
  // Parallel in time (scan.cu): the fused sweep + rollout of S segments of a.num_edges edges,
  // each from the boundary value function / state the scan computed; nullptr = not provided.
  int (*factor_solve_segments)(const FastArgs &, const double *Vb, const double *vb,
                               const double *xb, int S, cudaStream_t) = nullptr;
};

// nullptr when no specialised kernel exists for (n, m).
const FastPlan *select_fast_plan(int n, int m);
// CTA-per-problem DMMA kernels for large dimensions (riccati_cta.cu); nullptr when
// (n, m) is not instantiated.
const FastPlan *select_cta_plan(int n, int m);
// Reference-order register kernels for small chains (riccati_strict.cu): the smallest
// instantiated shape that holds (n, m), for chains padded to it; nullptr when none does.
const FastPlan *select_strict_plan(int n, int m);
const FastPlan *select_strict_tree_plan(int n, int m);

}  // namespace sipoc
