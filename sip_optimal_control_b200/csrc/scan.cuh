// Parallel-in-time front end (scan.cu): segment elements, boundary values and states.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "generic_kernels.cuh"
#include "profile.hpp"

namespace sipoc {

struct ScanArgs {
  LqrIn in;            // engine layout, uniform chain
  int N, M;            // state / control dimension
  int L, S;            // edges per segment, segments (L * S = horizon)
  int Sg;              // segments per group of the two-level boundary pass (S: one level)
  int64_t batch, ld;
  double *elems;       // [batch][S][3 N N + 2 N]  segment elements (A, C, J, b, eta)
  double *gelems;      // [batch][S / Sg][3 N N + 2 N]  group elements
  double *maps;        // [batch][S + S / Sg][N N + N]  x_end = T1 x_start + t2 (segments, groups)
  double *Vb, *vb, *xb;  // [S + 1][N N | N | N][ld]  boundary value functions and states
  int *seg_status;     // [S][ld]
  Profiler *prof;
};

bool scan_supports(int n, int m);
int64_t scan_elem_doubles(int n);
int64_t scan_map_doubles(int n);
int scan_group_size(int S);
// Kernels 1 and 2; returns the number of launches, -1 for an unsupported shape.
int launch_scan_front(const ScanArgs &a, cudaStream_t s);
// status[b] = first failure in post-order over the segments' sweep / scan statuses.
void launch_scan_status(const int *sweep_status, const int *seg_status, int S, int64_t batch,
                        int64_t ld, int *status, cudaStream_t s);

}  // namespace sipoc
