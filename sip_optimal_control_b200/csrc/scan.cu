// Parallel-in-time Riccati for long horizons and small batches (uniform chains).
//
// The reference walks the horizon serially (lqr.cpp:651, 738, 821); at N = 4 096 stages and
// 64 problems that leaves 8 warps on the whole GPU.  Here the horizon is cut into S segments
// of L edges and the recursion is restated on CONDITIONAL VALUE FUNCTIONS (the
// associative-scan elements of SURVEY.md section 7): an element (A, b, C, eta, J) of a run of
// edges a -> e represents
//     V(x_a; x_e) = max_lam  1/2 x_a' J x_a - x_a' eta - 1/2 lam' C lam - lam' (x_e - A x_a - b),
// one edge k (parent node k, child k + 1) has, from the inputs of lqr.hpp:76-89,
//     A = A_k - B_k R^-1 M_k',  b = c_{k+1} - B_k R^-1 r_k,  C = B_k R^-1 B_k' + diag(delta_{k+1}),
//     eta = -(q_k - M_k R^-1 r_k),  J = Q_k - M_k R^-1 M_k'
// (the delta-regularization of lqr.cpp:475-529 is a penalised dynamics slack, hence the
// diag(delta) in C; C > 0 always), and two consecutive runs combine as
//     Xi = (C1^-1 + J2)^-1
//     A  = A2 Xi C1^-1 A1                    b   = A2 Xi (C1^-1 b1 + eta2) + b2
//     C  = A2 Xi A2' + C2                    eta = A1' C1^-1 Xi (eta2 - J2 b1) + eta1
//     J  = A1' (C1^-1 - C1^-1 Xi C1^-1) A1 + J1.
// Combining with the terminal value (0, 0, 0, -v_e, V_e) gives the value function at a:
// V_a = J, v_a = -eta -- the reference's V, v at that node (checked to 1e-14 against the
// serial recursion in tests/test_oracle_scan.py's numpy restatement and on the device).
//
//   1. scan_segment_kernel   every (problem, segment) in parallel: the segment's element,
//                            absorbing its L edges one by one (L combines).
//   2. scan_boundary_kernel  per problem, S combines backward over the segment elements give
//                            (V, v) at every segment boundary and the boundary-to-boundary state
//                            maps x_e = T1 x_a + t2; S matrix-vector steps forward give x there.
//   3. riccati_backward_subwarp<..., SEGMENTED>  (riccati_fast.cu)  every (problem, segment)
//                            in parallel: the ordinary sweep + rollout of its L edges from the
//                            boundary data -- all of x, u, y, bit for bit the serial kernels'
//                            arithmetic inside a segment.
// Depth: 2 L + 2 S stage times instead of 2 N.  Needs R > 0 (the reference only needs
// R + B'WB > 0): a non-positive pivot of R is reported as G_FACTORIZATION_FAILURE.
//
// Mapping of 1 and 2: one WARP per problem (an element is five small dense objects; 32 lanes
// share each product), four problems per CTA so that global reads are whole 32-byte sectors
// of the batch-interleaved layout; all operands in shared memory.
#include "scan.cuh"

#include <cstdio>

namespace sipoc {
namespace {

constexpr int kWarps = 4;  // problems per CTA (32-byte sectors of the interleaved layout)

// ---- warp-level dense helpers on column-major N x N blocks in shared memory -------------
// C = alpha * op(A) * op(B) (+ C0).  TA / TB: use the transpose.
template <int N, bool TA, bool TB>
__device__ __forceinline__ void wmm(const double *A, const double *B, double *C, int lane,
                                    double alpha = 1.0, const double *C0 = nullptr) {
  for (int e = lane; e < N * N; e += 32) {
    const int i = e % N, j = e / N;
    double acc = 0.0, acc2 = 0.0;
#pragma unroll
    for (int k = 0; k < N; k += 2) {
      acc += (TA ? A[i * N + k] : A[k * N + i]) * (TB ? B[k * N + j] : B[j * N + k]);
      if (k + 1 < N)
        acc2 += (TA ? A[i * N + k + 1] : A[(k + 1) * N + i]) *
                (TB ? B[(k + 1) * N + j] : B[j * N + k + 1]);
    }
    C[e] = alpha * (acc + acc2) + (C0 != nullptr ? C0[e] : 0.0);
  }
}
// y = alpha * op(A) x (+ y0), lanes over rows.
template <int N, bool TA>
__device__ __forceinline__ void wmv(const double *A, const double *x, double *y, int lane,
                                    double alpha = 1.0, const double *y0 = nullptr) {
  if (lane < N) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) acc += (TA ? A[lane * N + k] : A[k * N + lane]) * x[k];
    y[lane] = alpha * acc + (y0 != nullptr ? y0[lane] : 0.0);
  }
}

// In-place inverse of a symmetric positive definite n x n block (n <= N runtime, column-major
// with leading dimension n): right-looking Cholesky shared by the lanes, the inverse of the
// factor one column per lane, then L^-T L^-1.  `work` holds n * n + n doubles.  Returns
// false (warp-uniform) on a pivot <= 0 -- Eigen::LLT's failure rule.
__device__ __forceinline__ bool spd_inverse(double *A, double *work, int n, int lane) {
  double *X = work, *dinv = work + n * n;
  bool ok = true;
  for (int j = 0; j < n; ++j) {
    const double piv = A[j * n + j];
    ok = ok && (piv > 0.0);
    const double s = rsqrt(piv);
    __syncwarp();
    if (lane == 0) dinv[j] = s;
    for (int i = j + lane; i < n; i += 32) A[j * n + i] *= s;  // column j of L (diag = sqrt)
    __syncwarp();
    const int rem = n - j - 1;
    for (int e = lane; e < rem * rem; e += 32) {  // trailing update, lower part
      const int c = j + 1 + e / rem, i = j + 1 + e % rem;
      if (i >= c) A[c * n + i] -= A[j * n + i] * A[j * n + c];
    }
    __syncwarp();
  }
  // X = L^-1, lane c owns column c
  for (int c = lane; c < n; c += 32) {
    for (int i = 0; i < c; ++i) X[c * n + i] = 0.0;
    X[c * n + c] = dinv[c];
    for (int i = c + 1; i < n; ++i) {
      double acc = 0.0;
      for (int k = c; k < i; ++k) acc += A[k * n + i] * X[c * n + k];
      X[c * n + i] = -acc * dinv[i];
    }
  }
  __syncwarp();
  // A^-1 = X' X
  for (int e = lane; e < n * n; e += 32) {
    const int i = e % n, j = e / n;
    double acc = 0.0;
    for (int k = (i > j ? i : j); k < n; ++k) acc += X[i * n + k] * X[j * n + k];
    A[e] = acc;
  }
  __syncwarp();
  return ok;
}

// One problem's working set in shared memory.
template <int N>
struct ScanSmem {
  static constexpr int NN = N * N;
  // accumulated element (the later run), incoming element (the earlier run), temporaries
  double A2[NN], C2[NN], J2[NN], b2[N], e2[N];
  double A1[NN], C1[NN], J1[NN], b1[N], e1[N];
  double Ci[NN], Xi[NN], CiA[NN], XA[NN], T[NN];
  double work[NN + N];
  double v0[N], v1[N], v2[N], v3[N];
};

// acc (2) <- combine(incoming (1), acc (2)).  Returns false if a factorization failed.
template <int N>
__device__ __forceinline__ bool combine(ScanSmem<N> &w, int lane, bool need_AC) {
  constexpr int NN = N * N;
  for (int e = lane; e < NN; e += 32) w.Ci[e] = w.C1[e];
  __syncwarp();
  bool ok = spd_inverse(w.Ci, w.work, N, lane);                    // C1^-1
  for (int e = lane; e < NN; e += 32) w.Xi[e] = w.Ci[e] + w.J2[e];
  __syncwarp();
  ok = spd_inverse(w.Xi, w.work, N, lane) && ok;                   // Xi = (C1^-1 + J2)^-1
  wmm<N, false, false>(w.Ci, w.A1, w.CiA, lane);                   // C1^-1 A1
  // v0 = C1^-1 b1 + eta2 ; v1 = eta2 - J2 b1
  wmv<N, false>(w.Ci, w.b1, w.v0, lane, 1.0, w.e2);
  wmv<N, false>(w.J2, w.b1, w.v1, lane, -1.0, w.e2);
  __syncwarp();
  wmm<N, false, false>(w.Xi, w.CiA, w.XA, lane);                   // X A1 = Xi C1^-1 A1
  wmv<N, false>(w.Xi, w.v0, w.v2, lane);                           // Xi (C1^-1 b1 + eta2)
  wmv<N, false>(w.Xi, w.v1, w.v3, lane);                           // Xi (eta2 - J2 b1)
  __syncwarp();
  // J = A1' C1^-1 A1 - (C1^-1 A1)' (X A1) + J1   (into T, then J2)
  wmm<N, true, false>(w.A1, w.CiA, w.T, lane, 1.0, w.J1);
  __syncwarp();
  wmm<N, true, false>(w.CiA, w.XA, w.J2, lane, -1.0, w.T);
  // eta = (C1^-1 A1)' Xi (eta2 - J2 b1) + eta1 ; b = A2 Xi (C1^-1 b1 + eta2) + b2
  wmv<N, true>(w.CiA, w.v3, w.v1, lane, 1.0, w.e1);
  if (need_AC) wmv<N, false>(w.A2, w.v2, w.v0, lane, 1.0, w.b2);
  __syncwarp();
  if (lane < N) {
    w.e2[lane] = w.v1[lane];
    if (need_AC) w.b2[lane] = w.v0[lane];
  }
  if (need_AC) {
    wmm<N, false, false>(w.A2, w.Xi, w.T, lane);                   // A2 Xi
    __syncwarp();
    wmm<N, false, true>(w.T, w.A2, w.Ci, lane, 1.0, w.C2);         // C = A2 Xi A2' + C2
    wmm<N, false, false>(w.A2, w.XA, w.CiA, lane);                 // A = A2 X A1
    __syncwarp();
    for (int e = lane; e < NN; e += 32) {
      w.C2[e] = w.Ci[e];
      w.A2[e] = w.CiA[e];
    }
  }
  __syncwarp();
  return ok;
}

// Element of edge k into slot 1 from the batch-interleaved inputs of problem b.
// `raw` (shared, >= 2 N N + 2 N M + M M + 3 N + 2 M + M M doubles) receives the stage first.
template <int N>
__device__ __forceinline__ bool load_edge_element(ScanSmem<N> &w, const LqrIn &in, int M, int k,
                                                  int64_t b, int64_t ld, int lane) {
  const size_t L = static_cast<size_t>(ld);
  auto G = [&](const double *p, size_t flat) { return __ldg(p + flat * L + b); };
  double *Bm = w.Ci, *Mm = w.Xi, *Ri = w.T, *BRi = w.CiA, *MRi = w.XA;  // scratch views
  const size_t kk = static_cast<size_t>(k);
  for (int e = lane; e < N * N; e += 32) {
    w.A1[e] = G(in.A, kk * N * N + e);
    const int i = e % N, j = e / N;  // symmetric Q from its lower triangle
    w.J1[e] = G(in.Q, kk * N * N + (i >= j ? j * N + i : i * N + j));
  }
  for (int e = lane; e < N * M; e += 32) {
    Bm[e] = G(in.B, kk * N * M + e);
    Mm[e] = G(in.M, kk * N * M + e);
  }
  for (int e = lane; e < M * M; e += 32) {
    const int i = e % M, j = e / M;
    Ri[e] = G(in.R, kk * M * M + (i >= j ? j * M + i : i * M + j));
  }
  if (lane < N) {
    w.b1[lane] = G(in.c, (kk + 1) * N + lane);
    w.e1[lane] = -G(in.q, kk * N + lane);
    w.v0[lane] = G(in.delta, (kk + 1) * N + lane);
  }
  if (lane < M) w.v1[lane] = G(in.r, kk * M + lane);
  __syncwarp();
  const bool ok = spd_inverse(Ri, w.work, M, lane);  // R^-1
  for (int e = lane; e < N * M; e += 32) {           // B R^-1 and M R^-1 (N x M)
    const int i = e % N, a = e / N;
    double sb = 0.0, sm = 0.0;
    for (int c = 0; c < M; ++c) {
      sb += Bm[c * N + i] * Ri[a * M + c];
      sm += Mm[c * N + i] * Ri[a * M + c];
    }
    BRi[e] = sb;
    MRi[e] = sm;
  }
  __syncwarp();
  for (int e = lane; e < N * N; e += 32) {
    const int i = e % N, j = e / N;
    double sa = 0.0, sc = 0.0, sj = 0.0;
    for (int a = 0; a < M; ++a) {
      sa += BRi[a * N + i] * Mm[a * N + j];   // B R^-1 M'
      sc += BRi[a * N + i] * Bm[a * N + j];   // B R^-1 B'
      sj += MRi[a * N + i] * Mm[a * N + j];   // M R^-1 M'
    }
    w.A1[e] -= sa;
    w.C1[e] = sc + (i == j ? w.v0[i] : 0.0);
    w.J1[e] -= sj;
  }
  if (lane < N) {
    double sb = 0.0, sm = 0.0;
    for (int a = 0; a < M; ++a) {
      sb += BRi[a * N + lane] * w.v1[a];
      sm += MRi[a * N + lane] * w.v1[a];
    }
    w.b1[lane] -= sb;   // c' - B R^-1 r
    w.e1[lane] += sm;   // -(q - M R^-1 r)
  }
  __syncwarp();
  return ok;
}

// ---- 1. segment elements ------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(32 * kWarps)
scan_segment_kernel(LqrIn in, int M, int L, int64_t batch, int64_t ld, double *elems,
                    int *seg_status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ScanSmem<N> &w = reinterpret_cast<ScanSmem<N> *>(smem_raw)[warp];
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
  const int seg = blockIdx.y, S = gridDim.y;
  if (b >= batch) return;
  constexpr int NN = N * N;
  for (int e = lane; e < NN; e += 32) {  // identity element
    w.A2[e] = (e % N == e / N) ? 1.0 : 0.0;
    w.C2[e] = 0.0;
    w.J2[e] = 0.0;
  }
  if (lane < N) w.b2[lane] = w.e2[lane] = 0.0;
  __syncwarp();
  bool ok = true, r_ok = true;
  for (int k = (seg + 1) * L - 1; k >= seg * L; --k) {
    r_ok = load_edge_element(w, in, M, k, b, ld, lane) && r_ok;
    ok = combine(w, lane, true) && ok;
  }
  double *out = elems + (static_cast<size_t>(b) * S + seg) * (3 * NN + 2 * N);
  for (int e = lane; e < NN; e += 32) {
    out[e] = w.A2[e];
    out[NN + e] = w.C2[e];
    out[2 * NN + e] = w.J2[e];
  }
  if (lane < N) {
    out[3 * NN + lane] = w.b2[lane];
    out[3 * NN + N + lane] = w.e2[lane];
  }
  if (lane == 0)
    seg_status[static_cast<size_t>(seg) * ld + b] =
        !r_ok ? SIPOC_FACTOR_G_FACTORIZATION_FAILURE
              : (!ok ? SIPOC_FACTOR_F_FACTORIZATION_FAILURE : SIPOC_FACTOR_SUCCESS);
}

// ---- 2. boundary values and states --------------------------------------------------------
// Vb [S + 1][N N][ld], vb / xb [S + 1][N][ld] in the engine layout (what the segmented sweep
// stages); T1 / t2 per (problem, segment) scratch, problem-major.
template <int N>
__global__ void __launch_bounds__(32 * kWarps)
scan_boundary_kernel(LqrIn in, int L, int S, int64_t batch, int64_t ld, const double *elems,
                     double *maps, double *Vb, double *vb, double *xb, int *seg_status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ScanSmem<N> &w = reinterpret_cast<ScanSmem<N> *>(smem_raw)[warp];
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
  if (b >= batch) return;
  constexpr int NN = N * N;
  const size_t Ld = static_cast<size_t>(ld);
  const size_t T = static_cast<size_t>(L) * S;
  // terminal value: V = Q_T, v = q_T  (lqr.cpp:658, 744)
  for (int e = lane; e < NN; e += 32) {
    const int i = e % N, j = e / N;
    w.J2[e] = __ldg(in.Q + (T * NN + (i >= j ? j * N + i : i * N + j)) * Ld + b);
    w.A2[e] = 0.0;
    w.C2[e] = 0.0;
  }
  if (lane < N) {
    w.e2[lane] = -__ldg(in.q + (T * N + lane) * Ld + b);
    w.b2[lane] = 0.0;
  }
  __syncwarp();
  bool ok = true;
  for (int seg = S; seg >= 0; --seg) {
    // publish (V, v) at boundary node seg * L
    for (int e = lane; e < NN; e += 32) Vb[(static_cast<size_t>(seg) * NN + e) * Ld + b] = w.J2[e];
    if (lane < N) vb[(static_cast<size_t>(seg) * N + lane) * Ld + b] = -w.e2[lane];
    if (seg == 0) break;
    const double *el = elems + (static_cast<size_t>(b) * S + (seg - 1)) * (3 * NN + 2 * N);
    for (int e = lane; e < NN; e += 32) {
      w.A1[e] = el[e];
      w.C1[e] = el[NN + e];
      w.J1[e] = el[2 * NN + e];
    }
    if (lane < N) {
      w.b1[lane] = el[3 * NN + lane];
      w.e1[lane] = el[3 * NN + N + lane];
    }
    __syncwarp();
    ok = combine(w, lane, false) && ok;
    // After combine: Xi = (C1^-1 + V_e)^-1, XA = Xi C1^-1 A1, v2 = Xi (C1^-1 b1 - v_e):
    // x_e = XA x_a + v2 (the state at the segment's end from the state at its start).
    double *mp = maps + (static_cast<size_t>(b) * S + (seg - 1)) * (NN + N);
    for (int e = lane; e < NN; e += 32) mp[e] = w.XA[e];
    if (lane < N) mp[NN + lane] = w.v2[lane];
    if (!ok && lane == 0 && seg_status[static_cast<size_t>(seg - 1) * Ld + b] == 0)
      seg_status[static_cast<size_t>(seg - 1) * Ld + b] = SIPOC_FACTOR_F_FACTORIZATION_FAILURE;
    __syncwarp();
  }
  // root: x_0 = -(I + D_0 V_0)^-1 (delta_0 o v_0 - c_0) = -(V_0 + D_0^-1)^-1 (v_0 - c_0 / delta_0)
  // (lqr.cpp:798-819; V_0 = J2, v_0 = -e2)
  if (lane < N) {
    const double d0 = __ldg(in.delta + static_cast<size_t>(lane) * Ld + b);
    const double c0 = __ldg(in.c + static_cast<size_t>(lane) * Ld + b);
    w.v0[lane] = 1.0 / d0;
    w.v1[lane] = -w.e2[lane] - c0 / d0;
  }
  __syncwarp();
  for (int e = lane; e < NN; e += 32) w.Xi[e] = w.J2[e] + (e % N == e / N ? w.v0[e % N] : 0.0);
  __syncwarp();
  spd_inverse(w.Xi, w.work, N, lane);
  wmv<N, false>(w.Xi, w.v1, w.v2, lane, -1.0);
  __syncwarp();
  if (lane < N) xb[static_cast<size_t>(lane) * Ld + b] = w.v2[lane];
  for (int seg = 0; seg < S; ++seg) {
    const double *mp = maps + (static_cast<size_t>(b) * S + seg) * (NN + N);
    if (lane < N) {
      double acc = mp[NN + lane];
#pragma unroll
      for (int k = 0; k < N; ++k) acc += mp[k * N + lane] * w.v2[k];
      w.v0[lane] = acc;
    }
    __syncwarp();
    if (lane < N) {
      w.v2[lane] = w.v0[lane];
      xb[(static_cast<size_t>(seg + 1) * N + lane) * Ld + b] = w.v0[lane];
    }
    __syncwarp();
  }
}

// The status of a problem is its first failure in post-order (lqr.cpp:696-700, 722-727): the
// failing segment with the largest index, sweep failures before scan failures of the same one.
__global__ void scan_status_kernel(const int *sweep_status, const int *seg_status, int S,
                                   int64_t batch, int64_t ld, int *status) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  int st = SIPOC_FACTOR_SUCCESS;
  for (int seg = S - 1; seg >= 0 && st == SIPOC_FACTOR_SUCCESS; --seg) {
    st = sweep_status[static_cast<size_t>(seg) * ld + b];
    if (st == SIPOC_FACTOR_SUCCESS) st = seg_status[static_cast<size_t>(seg) * ld + b];
  }
  status[b] = st;
}

template <int N>
int launch_front(const ScanArgs &a, cudaStream_t s) {
  const unsigned groups = static_cast<unsigned>((a.batch + kWarps - 1) / kWarps);
  const int bytes = static_cast<int>(sizeof(ScanSmem<N>)) * kWarps;
  auto k1 = scan_segment_kernel<N>;
  auto k2 = scan_boundary_kernel<N>;
  cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  {
    ProfScope ps(a.prof, "scan_segment_kernel", s);
    k1<<<dim3(groups, a.S), 32 * kWarps, bytes, s>>>(a.in, a.M, a.L, a.batch, a.ld, a.elems,
                                                     a.seg_status);
  }
  {
    ProfScope ps(a.prof, "scan_boundary_kernel", s);
    k2<<<groups, 32 * kWarps, bytes, s>>>(a.in, a.L, a.S, a.batch, a.ld, a.elems, a.maps, a.Vb,
                                          a.vb, a.xb, a.seg_status);
  }
  return 2;
}

}  // namespace

int64_t scan_elem_doubles(int n) { return 3LL * n * n + 2 * n; }
int64_t scan_map_doubles(int n) { return 1LL * n * n + n; }

int launch_scan_front(const ScanArgs &a, cudaStream_t s) {
  switch (a.N) {
    case 4: return launch_front<4>(a, s);
    case 6: return launch_front<6>(a, s);
    case 8: return launch_front<8>(a, s);
    case 12: return launch_front<12>(a, s);
    default: return -1;
  }
}

bool scan_supports(int n, int m) { return (n == 4 || n == 6 || n == 8 || n == 12) && m >= 1 && m <= n; }

void launch_scan_status(const int *sweep_status, const int *seg_status, int S, int64_t batch,
                        int64_t ld, int *status, cudaStream_t s) {
  scan_status_kernel<<<static_cast<unsigned>((batch + 127) / 128), 128, 0, s>>>(
      sweep_status, seg_status, S, batch, ld, status);
}

}  // namespace sipoc
