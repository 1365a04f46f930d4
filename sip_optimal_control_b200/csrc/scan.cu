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
// Mapping of 1 and 2: one WARP per problem (an element is five small dense objects), four
// problems per CTA so that global reads are whole 32-byte sectors of the batch-interleaved
// layout; all operands in shared memory.  Every N x N product runs on the FP64 tensor cores
// (`mma.sync.m8n8k4.f64`, SASS DMMA.8x8x4: one 8-byte shared-memory load per operand per 256
// FMA, against two loads per FMA for a lane-per-element product), the two inverses of a combine
// are symmetric sweeps (Gauss-Jordan on the lower triangle, three entries per lane in
// registers, one pivot per step -- positive pivots <=> the Cholesky pivots of Eigen::LLT).
// Kernel 1 stages edge k - 1 with cp.async while edge k is absorbed.
#include "scan.cuh"

#include <cstddef>
#include <cstdio>

namespace sipoc {
namespace {

constexpr int kWarps = 4;  // problems per CTA (32-byte sectors of the interleaved layout)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// D(8x8) += A(8x4) B(4x8).  Lane l (g = l / 4, t = l % 4) holds A(g, t), B(t, g) and
// D(g, 2t), D(g, 2t + 1).
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// N x N product with inner dimension K on the tensor cores: out(i, j) <- sum_k fa(i, k) fb(k, j)
// through fs(i, j, sum).  Every operand fragment is read before any result is stored (with a
// warp barrier between), so fs may overwrite an operand.
template <int N, int K, class FA, class FB, class FS>
__device__ __forceinline__ void tile_product(int lane, FA fa, FB fb, FS fs) {
  constexpr int RT = (N + 7) / 8, KS = (K + 3) / 4;
  const int g = lane >> 2, t = lane & 3;
  double af[RT][KS], bf[RT][KS];
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const int i = 8 * r + g, k = 4 * s + t;
      const bool in = (N % 8 == 0 || i < N) && (K % 4 == 0 || k < K);
      af[r][s] = in ? fa(i, k) : 0.0;
      bf[r][s] = in ? fb(k, i) : 0.0;
    }
  __syncwarp();
  // the RT x RT output tiles advance together, one k-step at a time: independent accumulators
  double acc[RT][RT][2];
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int c = 0; c < RT; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
#pragma unroll
  for (int s = 0; s < KS; ++s)
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < RT; ++c) dmma(acc[r][c], af[r][s], bf[c][s]);
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int c = 0; c < RT; ++c) {
      const int i = 8 * r + g, j = 8 * c + 2 * t;
      if (N % 8 == 0 || i < N) {
        if (N % 8 == 0 || j < N) fs(i, j, acc[r][c][0]);
        if (N % 8 == 0 || j + 1 < N) fs(i, j + 1, acc[r][c][1]);
      }
    }
}

// y = alpha * op(A) x (+ y0), lanes over rows.
template <int N, bool TA>
__device__ __forceinline__ void wmv(const double *A, const double *x, double *y, int lane,
                                    double alpha = 1.0, const double *y0 = nullptr) {
  if (lane < N) {
    double acc = 0.0, acc2 = 0.0;
#pragma unroll
    for (int k = 0; k < N; k += 2) {
      acc += (TA ? A[lane * N + k] : A[k * N + lane]) * x[k];
      if (k + 1 < N) acc2 += (TA ? A[lane * N + k + 1] : A[(k + 1) * N + lane]) * x[k + 1];
    }
    y[lane] = alpha * (acc + acc2) + (y0 != nullptr ? y0[lane] : 0.0);
  }
}

// 1 / d to about an ulp: MUFU.RCP64H seed and two Newton steps (no special cases: d is a
// pivot, and a pivot <= 0 is reported, not used).
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  return fma(x, e, x);
}

// In-place inverse of a symmetric positive definite NS x NS block (column-major, leading
// dimension NS, stored in full, 16-byte aligned) by symmetric sweeps on 2 x 2 pivot blocks
// J = {j, j + 1}:
//   A_JJ <- -P^-1;  A_iJ <- A_iJ P^-1;  A_ic <- A_ic - A_iJ P^-1 A_Jc   (i, c not in J),  P = A_JJ
// which leave -A^-1 after NS / 2 steps (half the dependent chain of one-pivot sweeps).  Both
// pivots of a block are positive (a_jj > 0 and det P > 0) exactly when the two Cholesky pivots
// are, so `all positive` is Eigen::LLT's success rule.  Lane i < NS keeps row i in registers
// (by symmetry it is stored contiguously as column i); the two pivot rows travel through `buf`
// (4 NS doubles, one half per step parity, so one warp barrier per step); the loop over the
// blocks is unrolled, every register index is static.  Returns false (warp-uniform) on a
// non-positive pivot.
// DUAL: lanes 16.. sweep a second block at A + NS NS with the buffer at buf + 4 NS.
template <int NS, bool DUAL = false>
__device__ __forceinline__ bool sweep_inverse(double *A, double *buf, int full_lane) {
  static_assert(NS % 2 == 0, "rows move as 16-byte pairs, pivots as 2 x 2 blocks");
  static_assert(!DUAL || NS <= 16, "two blocks share the warp");
  const int lane = DUAL ? full_lane & 15 : full_lane;
  if (DUAL && full_lane >= 16) {
    A += NS * NS;
    buf += 4 * NS;
  }
  const int i = lane < NS ? lane : 0;  // lanes beyond NS shadow row 0 and never store
  double a[NS];
  double2 *row = reinterpret_cast<double2 *>(A + i * NS);
#pragma unroll
  for (int c = 0; c < NS; c += 2) {
    const double2 v = row[c / 2];
    a[c] = v.x;
    a[c + 1] = v.y;
  }
  if (lane < 2) {
#pragma unroll
    for (int c = 0; c < NS; c += 2)
      reinterpret_cast<double2 *>(buf + lane * NS)[c / 2] = make_double2(a[c], a[c + 1]);
  }
  __syncwarp();
  bool ok = true;
#pragma unroll
  for (int jb = 0; jb < NS / 2; ++jb) {
    const int j = 2 * jb, k = j + 1;
    const double2 *uj = reinterpret_cast<const double2 *>(buf + (jb & 1) * 2 * NS);
    const double2 *uk = uj + NS / 2;
    double u[NS], w[NS];
#pragma unroll
    for (int c = 0; c < NS; c += 2) {
      const double2 v = uj[c / 2], z = uk[c / 2];
      u[c] = v.x;
      u[c + 1] = v.y;
      w[c] = z.x;
      w[c + 1] = z.y;
    }
    const double det = fma(u[j], w[k], -u[k] * u[k]);
    ok = ok && (u[j] > 0.0) && (det > 0.0);
    const double r = fast_rcp(det);
    const double p00 = w[k] * r, p01 = -u[k] * r, p11 = u[j] * r;  // P^-1
    if (lane == j || lane == k) {
      const double s0 = lane == j ? p00 : p01, s1 = lane == j ? p01 : p11;
#pragma unroll
      for (int c = 0; c < NS; ++c)
        a[c] = c == j ? -s0 : (c == k ? -s1 : fma(s0, u[c], s1 * w[c]));
    } else {
      const double g0 = fma(a[j], p00, a[k] * p01), g1 = fma(a[j], p01, a[k] * p11);
#pragma unroll
      for (int c = 0; c < NS; ++c)
        a[c] = c == j ? g0 : (c == k ? g1 : fma(-g0, u[c], fma(-g1, w[c], a[c])));
    }
    if (jb + 1 < NS / 2 && (lane == j + 2 || lane == j + 3)) {
      double2 *nx = reinterpret_cast<double2 *>(buf + (((jb + 1) & 1) * 2 + (lane - j - 2)) * NS);
#pragma unroll
      for (int c = 0; c < NS; c += 2) nx[c / 2] = make_double2(a[c], a[c + 1]);
    }
    __syncwarp();
  }
  if (lane < NS) {
#pragma unroll
    for (int c = 0; c < NS; c += 2) row[c / 2] = make_double2(-a[c], -a[c + 1]);
  }
  __syncwarp();
  return DUAL ? __all_sync(0xffffffffu, ok) : ok;
}

// Doubles of one element (A, C, J, b, eta) -- the order of the slots of ScanSmem and of the
// element arrays in global memory.
template <int N>
constexpr int kElem = 3 * N * N + 2 * N;

// One problem's working set in shared memory.
template <int N>
struct ScanSmem {
  static constexpr int NN = N * N;
  static constexpr int kTmp = NN > 8 * N ? NN : 8 * N;  // also holds two N x 4 build operands
  // accumulated element (the later run), incoming element (the earlier run), temporaries.
  // Regions are reused inside a combine: C1 is inverted in place (Ci), then receives
  // XA = Xi C1^-1 A1 once C1^-1 has been applied; T = A2 Xi lands on CiA once that is dead.
  double A2[NN], C2[NN], J2[NN], b2[N], e2[N];
  double A1[NN], C1[NN], J1[NN], b1[N], e1[N];
  double Xi[kTmp], CiA[kTmp];
  double v0[N], v1[N], v2[N], v3[N];
  double small[64];  // R, R + B' D^-1 B (4 x 4 each) and their sweep buffers
  __device__ __forceinline__ double *Ci() { return C1; }
  __device__ __forceinline__ double *XA() { return C1; }
  __device__ __forceinline__ double *T() { return CiA; }
};
// Elements move between global memory and a slot as one run of kElem doubles starting at
// A2 / A1, and the sweeps use v0..v3 as one 4 N buffer: the members must be contiguous.
template <int N>
constexpr bool scan_smem_layout_ok() {
  using S = ScanSmem<N>;
  constexpr size_t d = sizeof(double), NN = N * N;
  return offsetof(S, C2) == offsetof(S, A2) + NN * d && offsetof(S, J2) == offsetof(S, C2) + NN * d &&
         offsetof(S, b2) == offsetof(S, J2) + NN * d && offsetof(S, e2) == offsetof(S, b2) + N * d &&
         offsetof(S, A1) == offsetof(S, e2) + N * d && offsetof(S, C1) == offsetof(S, A1) + NN * d &&
         offsetof(S, J1) == offsetof(S, C1) + NN * d && offsetof(S, b1) == offsetof(S, J1) + NN * d &&
         offsetof(S, e1) == offsetof(S, b1) + N * d && offsetof(S, v1) == offsetof(S, v0) + N * d &&
         offsetof(S, v2) == offsetof(S, v1) + N * d && offsetof(S, v3) == offsetof(S, v2) + N * d &&
         sizeof(S) % 16 == 0 && offsetof(S, v0) % 16 == 0 && offsetof(S, small) % 16 == 0;
}
static_assert(scan_smem_layout_ok<6>() && scan_smem_layout_ok<8>() && scan_smem_layout_ok<12>(),
              "ScanSmem members are addressed as contiguous runs");

// acc (2) <- combine(incoming (1), acc (2)).  Returns false if a factorization failed.
// HAVE_CI: slot C1 already holds C1^-1 (the edge builder forms it by the Woodbury identity).
template <int N, bool NEED_AC, bool HAVE_CI = false>
__device__ __forceinline__ bool combine(ScanSmem<N> &w, int lane) {
  constexpr int NN = N * N;
  double *Ci = w.Ci(), *XA = w.XA(), *T = w.T();
  bool ok = true;
  if (!HAVE_CI) ok = sweep_inverse<N>(Ci, w.v0, lane);             // C1^-1, in place
  for (int e = lane; e < NN; e += 32) w.Xi[e] = Ci[e] + w.J2[e];
  __syncwarp();
  ok = sweep_inverse<N>(w.Xi, w.v0, lane) && ok;                  // Xi = (C1^-1 + J2)^-1
  wmv<N, false>(w.J2, w.b1, w.v1, lane, -1.0, w.e2);               // v1 = eta2 - J2 b1
  __syncwarp();
  tile_product<N, N>(lane, [&](int i, int k) { return Ci[k * N + i]; },
                     [&](int k, int j) { return w.A1[j * N + k]; },
                     [&](int i, int j, double v) { w.CiA[j * N + i] = v; });   // C1^-1 A1
  wmv<N, false>(Ci, w.b1, w.v0, lane, 1.0, w.e2);                  // v0 = C1^-1 b1 + eta2
  wmv<N, false>(w.Xi, w.v1, w.v3, lane);                           // Xi (eta2 - J2 b1)
  __syncwarp();
  tile_product<N, N>(lane, [&](int i, int k) { return w.Xi[k * N + i]; },
                     [&](int k, int j) { return w.CiA[j * N + k]; },
                     [&](int i, int j, double v) { XA[j * N + i] = v; });      // Xi C1^-1 A1
  wmv<N, false>(w.Xi, w.v0, w.v2, lane);                           // Xi (C1^-1 b1 + eta2)
  // eta = (C1^-1 A1)' Xi (eta2 - J2 b1) + eta1   (into v1, published below)
  wmv<N, true>(w.CiA, w.v3, w.v1, lane, 1.0, w.e1);
  __syncwarp();
  // J = J1 + A1' (C1^-1 A1) - (C1^-1 A1)' (Xi C1^-1 A1): one accumulation over 2 N terms
  tile_product<N, 2 * N>(
      lane, [&](int i, int k) { return k < N ? w.A1[i * N + k] : -w.CiA[i * N + k - N]; },
      [&](int k, int j) { return k < N ? w.CiA[j * N + k] : XA[j * N + k - N]; },
      [&](int i, int j, double v) { w.J2[j * N + i] = v + w.J1[j * N + i]; });
  if (lane < N) w.e2[lane] = w.v1[lane];
  if (NEED_AC) {
    wmv<N, false>(w.A2, w.v2, w.v0, lane, 1.0, w.b2);               // b = A2 Xi (...) + b2
    tile_product<N, N>(lane, [&](int i, int k) { return w.A2[k * N + i]; },
                       [&](int k, int j) { return w.Xi[j * N + k]; },
                       [&](int i, int j, double v) { T[j * N + i] = v; });     // A2 Xi
    __syncwarp();
    if (lane < N) w.b2[lane] = w.v0[lane];
    tile_product<N, N>(lane, [&](int i, int k) { return T[k * N + i]; },
                       [&](int k, int j) { return w.A2[k * N + j]; },
                       [&](int i, int j, double v) { w.C2[j * N + i] += v; }); // C = A2 Xi A2' + C2
    __syncwarp();
    tile_product<N, N>(lane, [&](int i, int k) { return w.A2[k * N + i]; },
                       [&](int k, int j) { return XA[j * N + k]; },
                       [&](int i, int j, double v) { w.A2[j * N + i] = v; });  // A = A2 Xi C1^-1 A1
  }
  __syncwarp();
  return ok;
}

// Flat offsets of one edge's operands in the staging buffer (doubles per problem).
struct EdgeMap {
  int A, B, Q, M, R, c, q, d, r, total;
  __device__ __host__ EdgeMap(int n, int m) {
    A = 0;
    B = A + n * n;
    Q = B + n * m;
    M = Q + n * n;
    R = M + n * m;
    c = R + m * m;
    q = c + n;
    d = q + n;
    r = d + n;
    total = r + m;
  }
};

// Edge k of the four problems of the CTA -> raw[flat][4] (16-byte copies, whole sectors).
__device__ __forceinline__ void stage_edge(double *raw, const LqrIn &in, const EdgeMap &mp, int n,
                                           int m, int k, int64_t b0, int64_t ld) {
  const size_t L = static_cast<size_t>(ld), kk = static_cast<size_t>(k);
  const int half = threadIdx.x & 1;
  for (int f = threadIdx.x >> 1; f < mp.total; f += (32 * kWarps) >> 1) {
    const double *src;
    if (f < mp.B) src = in.A + (kk * n * n + (f - mp.A)) * L;
    else if (f < mp.Q) src = in.B + (kk * n * m + (f - mp.B)) * L;
    else if (f < mp.M) src = in.Q + (kk * n * n + (f - mp.Q)) * L;
    else if (f < mp.R) src = in.M + (kk * n * m + (f - mp.M)) * L;
    else if (f < mp.c) src = in.R + (kk * m * m + (f - mp.R)) * L;
    else if (f < mp.q) src = in.c + ((kk + 1) * n + (f - mp.c)) * L;
    else if (f < mp.d) src = in.q + (kk * n + (f - mp.q)) * L;
    else if (f < mp.r) src = in.delta + ((kk + 1) * n + (f - mp.d)) * L;
    else src = in.r + (kk * m + (f - mp.r)) * L;
    cp_async16(raw + f * kWarps + 2 * half, src + b0 + 2 * half);
  }
}

// Element of the staged edge into slot 1 (raw is this warp's column of the staging buffer:
// entry f at raw[f * kWarps]).  FIRST: slot C1 receives C = B R^-1 B' + D (the run starts as
// this element); otherwise it receives C^-1 by the Woodbury identity
//   C^-1 = D^-1 - D^-1 B (R + B' D^-1 B)^-1 B' D^-1
// -- a second 4 x 4 inverse (swept beside R^-1 by the upper half of the warp) instead of an
// N x N one.
template <int N, bool FIRST>
__device__ __forceinline__ bool build_edge_element(ScanSmem<N> &w, const double *raw,
                                                   const EdgeMap &mp, int M, int lane) {
  auto R = [&](int f) { return raw[f * kWarps]; };
  double *Ri = w.small, *Rt = w.small + 16, *buf = w.small + 32;
  double *BRi = w.Xi, *MRi = w.Xi + 4 * N, *Bd = w.CiA, *BG = w.CiA + 4 * N;  // N x 4 each
  if (!FIRST) {
    for (int e = lane; e < N * M; e += 32) Bd[e] = R(mp.B + e) / R(mp.d + e % N);  // D^-1 B
    __syncwarp();
  }
  {  // symmetric R from its lower triangle, padded to 4 x 4 with the identity; R + B' D^-1 B
    const int l = lane & 15, i = l & 3, j = l >> 2;
    const bool live = i < M && j < M;
    double v = live ? R(mp.R + (i >= j ? j * M + i : i * M + j)) : (i == j ? 1.0 : 0.0);
    if (lane >= 16 && !FIRST && live) {
      double acc = 0.0, acc2 = 0.0;
#pragma unroll
      for (int k = 0; k < N; k += 2) {
        acc += R(mp.B + i * N + k) * Bd[j * N + k];
        acc2 += R(mp.B + i * N + k + 1) * Bd[j * N + k + 1];
      }
      v += acc + acc2;
    }
    w.small[lane] = v;
  }
  __syncwarp();
  const bool ok = sweep_inverse<4, !FIRST>(Ri, buf, lane);  // R^-1 (and (R + B' D^-1 B)^-1)
  for (int e = lane; e < N * M; e += 32) {                  // B R^-1, M R^-1, D^-1 B G~ (N x M)
    const int i = e % N, a = e / N;
    double sb = 0.0, sm = 0.0, sg = 0.0;
    for (int c = 0; c < M; ++c) {
      sb += R(mp.B + c * N + i) * Ri[a * 4 + c];
      sm += R(mp.M + c * N + i) * Ri[a * 4 + c];
      if (!FIRST) sg += Bd[c * N + i] * Rt[a * 4 + c];
    }
    BRi[e] = sb;
    MRi[e] = sm;
    if (!FIRST) BG[e] = sg;
  }
  __syncwarp();
  // A = A_k - B R^-1 M',  J = Q_k - M R^-1 M',  C = B R^-1 B' + diag(delta')  (inner dim M <= 4)
  tile_product<N, 4>(lane, [&](int i, int k) { return k < M ? BRi[k * N + i] : 0.0; },
                     [&](int k, int j) { return k < M ? R(mp.M + k * N + j) : 0.0; },
                     [&](int i, int j, double v) { w.A1[j * N + i] = R(mp.A + j * N + i) - v; });
  tile_product<N, 4>(lane, [&](int i, int k) { return k < M ? MRi[k * N + i] : 0.0; },
                     [&](int k, int j) { return k < M ? R(mp.M + k * N + j) : 0.0; },
                     [&](int i, int j, double v) {
                       w.J1[j * N + i] = R(mp.Q + (i >= j ? j * N + i : i * N + j)) - v;
                     });
  if (FIRST) {
    tile_product<N, 4>(lane, [&](int i, int k) { return k < M ? BRi[k * N + i] : 0.0; },
                       [&](int k, int j) { return k < M ? R(mp.B + k * N + j) : 0.0; },
                       [&](int i, int j, double v) {
                         w.C1[j * N + i] = v + (i == j ? R(mp.d + i) : 0.0);
                       });
  } else {
    tile_product<N, 4>(lane, [&](int i, int k) { return k < M ? BG[k * N + i] : 0.0; },
                       [&](int k, int j) { return k < M ? Bd[k * N + j] : 0.0; },
                       [&](int i, int j, double v) {
                         w.C1[j * N + i] = (i == j ? 1.0 / R(mp.d + i) : 0.0) - v;
                       });
  }
  if (lane < N) {
    double sb = 0.0, sm = 0.0;
    for (int a = 0; a < M; ++a) {
      sb += BRi[a * N + lane] * R(mp.r + a);
      sm += MRi[a * N + lane] * R(mp.r + a);
    }
    w.b1[lane] = R(mp.c + lane) - sb;   // c' - B R^-1 r
    w.e1[lane] = sm - R(mp.q + lane);   // -(q - M R^-1 r)
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
  double *raw = reinterpret_cast<double *>(smem_raw + sizeof(ScanSmem<N>) * kWarps);
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * kWarps, b = b0 + warp;
  const int seg = blockIdx.y, S = gridDim.y;
  const bool active = b < batch;  // idle warps still take part in the CTA barriers
  const EdgeMap mp(N, M);
  const int k_last = (seg + 1) * L - 1, k_first = seg * L;
  stage_edge(raw, in, mp, N, M, k_last, b0, ld);
  cp_async_commit();
  bool ok = true, r_ok = true;
  for (int k = k_last; k >= k_first; --k) {
    cp_async_wait_all();
    __syncthreads();
    if (active)
      r_ok = (k == k_last ? build_edge_element<N, true>(w, raw + warp, mp, M, lane)
                          : build_edge_element<N, false>(w, raw + warp, mp, M, lane)) && r_ok;
    __syncthreads();  // every warp is done with the staging buffer
    if (k > k_first) stage_edge(raw, in, mp, N, M, k - 1, b0, ld);
    cp_async_commit();
    if (!active) continue;
    if (k == k_last) {  // the run starts as the last edge's own element
      for (int e = lane; e < kElem<N>; e += 32) w.A2[e] = w.A1[e];
      __syncwarp();
    } else {
      ok = combine<N, true, true>(w, lane) && ok;
    }
  }
  if (!active) return;
  double *out = elems + (static_cast<size_t>(b) * S + seg) * kElem<N>;
  for (int e = lane; e < kElem<N>; e += 32) out[e] = w.A2[e];
  if (lane == 0)
    seg_status[static_cast<size_t>(seg) * ld + b] =
        !r_ok ? SIPOC_FACTOR_G_FACTORIZATION_FAILURE
              : (!ok ? SIPOC_FACTOR_F_FACTORIZATION_FAILURE : SIPOC_FACTOR_SUCCESS);
}

// ---- 2. boundary values and states --------------------------------------------------------
// Two levels so that the dependent chain is 2 Sg + G combines instead of S = Sg G:
//   2a. scan_group_kernel   every (problem, group of Sg segments): the group's element.
//   2b. scan_chain_kernel   (top) per problem, the chain over the G group elements from the
//       terminal node: (V, v) and x at the group boundaries.
//   2c. scan_chain_kernel   every (problem, group): the chain over its Sg segment elements from
//       the group's end boundary: (V, v) and x at the boundaries inside the group.
// With G == 1 only 2b runs, on the segment elements.
template <int N>
__global__ void __launch_bounds__(32 * kWarps)
scan_group_kernel(int Sg, int G, int64_t batch, int64_t ld, const double *elems, double *gelems,
                  int *seg_status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ScanSmem<N> &w = reinterpret_cast<ScanSmem<N> *>(smem_raw)[warp];
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
  const int g = blockIdx.y;
  if (b >= batch) return;
  constexpr int EL = kElem<N>, PER = (EL + 31) / 32;
  const double *el = elems + (static_cast<size_t>(b) * G * Sg + static_cast<size_t>(g + 1) * Sg - 1) * EL;
  for (int e = lane; e < EL; e += 32) w.A2[e] = el[e];
  bool ok = true;
  double pre[PER];
  for (int t = Sg - 2; t >= 0; --t) {
    el -= EL;
#pragma unroll
    for (int q = 0; q < PER; ++q) pre[q] = lane + 32 * q < EL ? el[lane + 32 * q] : 0.0;
#pragma unroll
    for (int q = 0; q < PER; ++q)
      if (lane + 32 * q < EL) w.A1[lane + 32 * q] = pre[q];
    __syncwarp();
    ok = combine<N, true>(w, lane) && ok;
  }
  __syncwarp();
  double *out = gelems + (static_cast<size_t>(b) * G + g) * EL;
  for (int e = lane; e < EL; e += 32) out[e] = w.A2[e];
  const size_t first = static_cast<size_t>(g) * Sg * ld + b;
  if (!ok && lane == 0 && seg_status[first] == 0) seg_status[first] = SIPOC_FACTOR_F_FACTORIZATION_FAILURE;
}

// Vb [S + 1][N N][ld], vb / xb [S + 1][N][ld] in the engine layout (what the segmented sweep
// stages).  Chain h of problem b runs over elements el[b][h cnt .. (h + 1) cnt), each spanning
// `span` segments; maps (x_end = T1 x_start + t2 per element) are scratch, problem-major.
template <int N, bool TOP>
__global__ void __launch_bounds__(32 * kWarps)
scan_chain_kernel(LqrIn in, int L, int cnt, int span, int H, int64_t batch, int64_t ld,
                  const double *elems, double *maps, double *Vb, double *vb, double *xb,
                  int *seg_status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ScanSmem<N> &w = reinterpret_cast<ScanSmem<N> *>(smem_raw)[warp];
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
  const int h = blockIdx.y;
  if (b >= batch) return;
  constexpr int NN = N * N, EL = kElem<N>, PER = (EL + 31) / 32;
  const size_t Ld = static_cast<size_t>(ld);
  const size_t e0 = static_cast<size_t>(h) * cnt;        // first element of the chain
  const size_t bnd_end = (e0 + cnt) * span;              // boundary index of the chain's end
  if (TOP) {
    // terminal value: V = Q_T, v = q_T  (lqr.cpp:658, 744)
    const size_t T = bnd_end * L;  // the chain of the top level ends at the last node
    for (int e = lane; e < NN; e += 32) {
      const int i = e % N, j = e / N;
      w.J2[e] = __ldg(in.Q + (T * NN + (i >= j ? j * N + i : i * N + j)) * Ld + b);
    }
    if (lane < N) w.e2[lane] = -__ldg(in.q + (T * N + lane) * Ld + b);
  } else {
    for (int e = lane; e < NN; e += 32) w.J2[e] = Vb[(bnd_end * NN + e) * Ld + b];
    if (lane < N) w.e2[lane] = -vb[(bnd_end * N + lane) * Ld + b];
  }
  const double *el = elems + (static_cast<size_t>(b) * H * cnt + e0 + cnt - 1) * EL;
  double *mp = maps + (static_cast<size_t>(b) * H * cnt + e0 + cnt - 1) * (NN + N);
  double pre[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) pre[q] = lane + 32 * q < EL ? el[lane + 32 * q] : 0.0;
  __syncwarp();
  bool ok = true;
  for (int t = cnt; t >= 0; --t) {
    // publish (V, v) at boundary (e0 + t) span -- the chain's ends are the parent level's
    if (TOP || (t > 0 && t < cnt)) {
      const size_t bnd = (e0 + t) * span;
      for (int e = lane; e < NN; e += 32) Vb[(bnd * NN + e) * Ld + b] = w.J2[e];
      if (lane < N) vb[(bnd * N + lane) * Ld + b] = -w.e2[lane];
    }
    if (t == 0) break;
#pragma unroll
    for (int q = 0; q < PER; ++q)
      if (lane + 32 * q < EL) w.A1[lane + 32 * q] = pre[q];
    if (t > 1) {  // the next element loads under the combine
      el -= EL;
#pragma unroll
      for (int q = 0; q < PER; ++q) pre[q] = lane + 32 * q < EL ? el[lane + 32 * q] : 0.0;
    }
    __syncwarp();
    ok = combine<N, false>(w, lane) && ok;
    // After combine: Xi = (C1^-1 + V_e)^-1, XA = Xi C1^-1 A1, v2 = Xi (C1^-1 b1 - v_e):
    // x_e = XA x_a + v2 (the state at the element's end from the state at its start).
    for (int e = lane; e < NN; e += 32) mp[e] = w.XA()[e];
    if (lane < N) mp[NN + lane] = w.v2[lane];
    mp -= NN + N;
    const size_t first = (e0 + t - 1) * span * Ld + b;
    if (!ok && lane == 0 && seg_status[first] == 0)
      seg_status[first] = SIPOC_FACTOR_F_FACTORIZATION_FAILURE;
    __syncwarp();
  }
  mp += NN + N;  // map of the chain's first element
  if (TOP) {
    // root: x_0 = -(I + D_0 V_0)^-1 (delta_0 o v_0 - c_0) = -(V_0 + D_0^-1)^-1 (v_0 - c_0 / delta_0)
    // (lqr.cpp:798-819; V_0 = J2, v_0 = -e2)
    if (lane < N) {
      const double d0 = __ldg(in.delta + static_cast<size_t>(lane) * Ld + b);
      const double c0 = __ldg(in.c + static_cast<size_t>(lane) * Ld + b);
      w.b1[lane] = 1.0 / d0;
      w.e1[lane] = -w.e2[lane] - c0 / d0;
    }
    __syncwarp();
    for (int e = lane; e < NN; e += 32) w.Xi[e] = w.J2[e] + (e % N == e / N ? w.b1[e % N] : 0.0);
    __syncwarp();
    sweep_inverse<N>(w.Xi, w.v0, lane);
    wmv<N, false>(w.Xi, w.e1, w.v0, lane, -1.0);
    __syncwarp();
    if (lane < N) xb[static_cast<size_t>(lane) * Ld + b] = w.v0[lane];
  } else {
    if (lane < N) w.v0[lane] = xb[(e0 * span * N + lane) * Ld + b];
  }
  __syncwarp();
  for (int t = 0; t < cnt; ++t, mp += NN + N) {
    if (lane < N) {
      double acc = mp[NN + lane];
#pragma unroll
      for (int k = 0; k < N; ++k) acc += mp[k * N + lane] * w.v0[k];
      w.v1[lane] = acc;
    }
    __syncwarp();
    if (lane < N) {
      w.v0[lane] = w.v1[lane];
      if (TOP || t + 1 < cnt) xb[((e0 + t + 1) * span * N + lane) * Ld + b] = w.v1[lane];
    }
    __syncwarp();
  }
}

// The status of a problem is its first failure in post-order (lqr.cpp:696-700, 722-727): the
// failing segment with the largest index, sweep failures before scan failures of the same one.
// One warp per problem, lanes over the segments.
__global__ void scan_status_kernel(const int *sweep_status, const int *seg_status, int S,
                                   int64_t batch, int64_t ld, int *status) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= batch) return;
  int key = -1, code = SIPOC_FACTOR_SUCCESS;  // key = 2 seg + (1 for a sweep failure)
  for (int seg = lane; seg < S; seg += 32) {
    const int sw = sweep_status[static_cast<size_t>(seg) * ld + b];
    const int sc = seg_status[static_cast<size_t>(seg) * ld + b];
    if (sw != SIPOC_FACTOR_SUCCESS) {
      key = 2 * seg + 1;
      code = sw;
    } else if (sc != SIPOC_FACTOR_SUCCESS) {
      key = 2 * seg;
      code = sc;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int k2 = __shfl_xor_sync(0xffffffffu, key, o), c2 = __shfl_xor_sync(0xffffffffu, code, o);
    if (k2 > key) {
      key = k2;
      code = c2;
    }
  }
  if (lane == 0) status[b] = code;
}

template <int N>
int launch_front(const ScanArgs &a, cudaStream_t s) {
  const unsigned groups = static_cast<unsigned>((a.batch + kWarps - 1) / kWarps);
  const int bytes = static_cast<int>(sizeof(ScanSmem<N>)) * kWarps;
  const int bytes1 = bytes + EdgeMap(N, a.M).total * kWarps * static_cast<int>(sizeof(double));
  auto k1 = scan_segment_kernel<N>;
  auto k2a = scan_group_kernel<N>;
  auto k2b = scan_chain_kernel<N, true>;
  auto k2c = scan_chain_kernel<N, false>;
  cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes1);
  cudaFuncSetAttribute(k2a, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(k2b, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(k2c, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  {
    ProfScope ps(a.prof, "scan_segment_kernel", s);
    k1<<<dim3(groups, a.S), 32 * kWarps, bytes1, s>>>(a.in, a.M, a.L, a.batch, a.ld, a.elems,
                                                      a.seg_status);
  }
  const int Sg = a.Sg, G = a.S / a.Sg;
  if (G == 1) {
    ProfScope ps(a.prof, "scan_chain_kernel", s);
    k2b<<<groups, 32 * kWarps, bytes, s>>>(a.in, a.L, a.S, 1, 1, a.batch, a.ld, a.elems, a.maps, a.Vb,
                                           a.vb, a.xb, a.seg_status);
    return 2;
  }
  double *gmaps = a.maps + a.batch * a.S * scan_map_doubles(N);
  {
    ProfScope ps(a.prof, "scan_group_kernel", s);
    k2a<<<dim3(groups, G), 32 * kWarps, bytes, s>>>(Sg, G, a.batch, a.ld, a.elems, a.gelems,
                                                    a.seg_status);
  }
  {
    ProfScope ps(a.prof, "scan_chain_kernel(groups)", s);
    k2b<<<groups, 32 * kWarps, bytes, s>>>(a.in, a.L, G, Sg, 1, a.batch, a.ld, a.gelems, gmaps, a.Vb,
                                           a.vb, a.xb, a.seg_status);
  }
  {
    ProfScope ps(a.prof, "scan_chain_kernel(segments)", s);
    k2c<<<dim3(groups, G), 32 * kWarps, bytes, s>>>(a.in, a.L, Sg, 1, G, a.batch, a.ld, a.elems, a.maps,
                                                    a.Vb, a.vb, a.xb, a.seg_status);
  }
  return 4;
}

}  // namespace

int64_t scan_elem_doubles(int n) { return 3LL * n * n + 2 * n; }

// Segments per group of the two-level boundary pass: the divisor of S that minimises the
// dependent chain 2 Sg + S / Sg (S itself -- one level -- for short chains).
int scan_group_size(int S) {
  if (S <= 12) return S;
  int best = S;
  for (int sg = 2; sg < S; ++sg)
    if (S % sg == 0 && 2 * sg + S / sg < (best == S ? S : 2 * best + S / best)) best = sg;
  return best;
}
int64_t scan_map_doubles(int n) { return 1LL * n * n + n; }

int launch_scan_front(const ScanArgs &a, cudaStream_t s) {
  switch (a.N) {
    case 6: return launch_front<6>(a, s);
    case 8: return launch_front<8>(a, s);
    case 12: return launch_front<12>(a, s);
    default: return -1;
  }
}

bool scan_supports(int n, int m) { return (n == 6 || n == 8 || n == 12) && m >= 1 && m <= 4; }

void launch_scan_status(const int *sweep_status, const int *seg_status, int S, int64_t batch,
                        int64_t ld, int *status, cudaStream_t s) {
  scan_status_kernel<<<static_cast<unsigned>((batch + 3) / 4), 128, 0, s>>>(
      sweep_status, seg_status, S, batch, ld, status);
}

}  // namespace sipoc
