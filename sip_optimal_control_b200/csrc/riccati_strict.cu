// Reference-order kernels for small uniform chains: the statement order of lqr.cpp
// (factor_with_status :645-731, solve :735-871) kept operation for operation, as in the
// generic kernels, but with compile-time dims, every intermediate in registers and the
// stage operands staged through shared memory by cp.async a few stages ahead.
//
// They exist for chains whose dims change from stage to stage (config 4): such a chain
// is padded to the uniform shape (N, M) by pad_chain_kernel with states / controls that
// are decoupled from the real ones (identity on the padded diagonal of Q and R, delta = 1,
// zeros elsewhere).  The padding sits at the high end of every index range and enters
// every sum as an exact zero, so the real entries go through exactly the additions,
// multiplications and square roots the generic kernels perform on the unpadded problem,
// in the same order.  The one departure: the divisions by the diagonal of a Cholesky
// factor (three quarters of the generic kernels' instructions at these sizes) are
// multiplications by its reciprocal, computed once per factor -- at most an ulp per
// quotient.  That keeps the 1e-9 parity bar on ill-conditioned regularization (r2 up to
// 1e9: 8e-14 from the oracle, the generic kernels 5e-14) where the reordered
// shape-specialised kernels drift to 1e-5.
//
// One thread per problem, one warp per block.
#include <cstdio>

#include "riccati_fast.cuh"

namespace sipoc {
namespace {

__device__ __forceinline__ void cp_async8(double *smem, const double *gmem) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int PENDING>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING) : "memory");
}

constexpr int kBuf = 3;  // stages in flight

// Per-problem element counts of the kept factorization (interleaved [flat * ld + problem]).
template <int N, int M>
struct StrictSizes {
  // per node: V, F factor (N x N each), sqrt(delta), 1 / sqrt(delta); per edge: W, K, G factor
  static constexpr int kNode = 2 * N * N + 2 * N;
  static __host__ __device__ constexpr int64_t oV(int) { return 0; }
  static __host__ __device__ constexpr int64_t oF(int T) { return int64_t(T + 1) * N * N; }
  static __host__ __device__ constexpr int64_t oSd(int T) { return 2 * int64_t(T + 1) * N * N; }
  static __host__ __device__ constexpr int64_t oSdi(int T) { return oSd(T) + int64_t(T + 1) * N; }
  static __host__ __device__ constexpr int64_t oW(int T) { return int64_t(T + 1) * kNode; }
  static __host__ __device__ constexpr int64_t oK(int T) { return oW(T) + int64_t(T) * N * N; }
  static __host__ __device__ constexpr int64_t oG(int T) { return oK(T) + int64_t(T) * M * N; }
  static __host__ __device__ constexpr int64_t store(int T) { return oG(T) + int64_t(T) * M * M; }
  static __host__ __device__ constexpr int64_t ov(int) { return 0; }
  static __host__ __device__ constexpr int64_t ok(int T) { return int64_t(T + 1) * N; }
  static __host__ __device__ constexpr int64_t scratch(int T) { return ok(T) + int64_t(T) * M; }
};

// Unblocked left-looking lower Cholesky, column-major n x n in registers (chol_lower of
// generic_kernels.cu, Eigen LLT semantics: a pivot <= 0 fails).  On return the strict lower
// triangle holds L and the diagonal holds 1 / L(k, k).
template <int n>
__device__ __forceinline__ bool chol_lower(double (&a)[n * n]) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < n; ++k) {
    double x = a[k + k * n];
#pragma unroll
    for (int j = 0; j < k; ++j) {
      const double l = a[k + j * n];
      x -= l * l;
    }
    ok = ok && (x > 0.0);
    x = sqrt(x);
    // One division per pivot: the column is scaled by the reciprocal, and the reciprocal --
    // not the pivot -- is what is kept on the diagonal (the diagonal of a factor is only
    // ever divided by, here and in chol_solve; the kept factorization is private to
    // these kernels).
    const double rx = 1.0 / x;
    a[k + k * n] = rx;
#pragma unroll
    for (int i = k + 1; i < n; ++i) {
      double s = a[i + k * n];
#pragma unroll
      for (int j = 0; j < k; ++j) s -= a[i + j * n] * a[k + j * n];
      a[i + k * n] = s * rx;
    }
  }
  return ok;
}

// L X = B, then L' X = B, in place, column by column (chol_solve of generic_kernels.cu);
// `l` as chol_lower leaves it (reciprocal diagonal).
template <int n, int nrhs>
__device__ __forceinline__ void chol_solve(const double (&l)[n * n], double (&b)[n * nrhs]) {
  // The 2 n nrhs divisions by the n diagonal entries are multiplications by their
  // reciprocals, which chol_lower left on the diagonal (the one place these kernels
  // depart from the generic kernels' operations: at most an ulp per quotient).
#pragma unroll
  for (int c = 0; c < nrhs; ++c) {
#pragma unroll
    for (int i = 0; i < n; ++i) {
      double s = b[c * n + i];
#pragma unroll
      for (int j = 0; j < i; ++j) s -= l[i + j * n] * b[c * n + j];
      b[c * n + i] = s * l[i + i * n];
    }
#pragma unroll
    for (int i = n - 1; i >= 0; --i) {
      double s = b[c * n + i];
#pragma unroll
      for (int j = i + 1; j < n; ++j) s -= l[j + i * n] * b[c * n + j];
      b[c * n + i] = s * l[i + i * n];
    }
  }
}

// ===========================================================================
// factor_with_status (lqr.cpp:645-731) on a chain: nodes T .. 0.
// ===========================================================================
template <int N, int M>
struct FactorRows {
  static constexpr int rQ = 0, rD = rQ + N * N, rM = rD + N, rR = rM + N * M, rA = rR + M * M,
                       rB = rA + N * N, kRows = rB + N * M;
  static constexpr int kBytes = kBuf * kRows * 32 * int(sizeof(double));
};

template <int N, int M>
__global__ void __launch_bounds__(32)
strict_factor_thread(LqrIn in, int *status_out, double *store, int64_t batch, int64_t ld, int T) {
  using Z = StrictSizes<N, M>;
  using R = FactorRows<N, M>;
  extern __shared__ __align__(16) double sm_strict[];
  const int lane = threadIdx.x;
  const int64_t b_raw = static_cast<int64_t>(blockIdx.x) * 32 + lane;
  const bool valid = b_raw < batch;
  const int64_t b = valid ? b_raw : batch - 1;
  const size_t L_ = static_cast<size_t>(ld);
  double *st_base = store + b;
  auto put = [&](int64_t flat, double v) {
    if (valid) __stcs(st_base + static_cast<size_t>(flat) * L_, v);
  };

  // operands of node k and of its child edge k (k < T)
  auto fetch = [&](int k, int buf) {
    double *dst = sm_strict + static_cast<size_t>(buf) * R::kRows * 32 + lane;
    const size_t kk = static_cast<size_t>(k);
#pragma unroll
    for (int t = 0; t < N * N; ++t) cp_async8(dst + (R::rQ + t) * 32, in.Q + (kk * N * N + t) * L_ + b);
#pragma unroll
    for (int i = 0; i < N; ++i) cp_async8(dst + (R::rD + i) * 32, in.delta + (kk * N + i) * L_ + b);
    if (k < T) {
#pragma unroll
      for (int t = 0; t < N * M; ++t) {
        cp_async8(dst + (R::rM + t) * 32, in.M + (kk * N * M + t) * L_ + b);
        cp_async8(dst + (R::rB + t) * 32, in.B + (kk * N * M + t) * L_ + b);
      }
#pragma unroll
      for (int t = 0; t < M * M; ++t) cp_async8(dst + (R::rR + t) * 32, in.R + (kk * M * M + t) * L_ + b);
#pragma unroll
      for (int t = 0; t < N * N; ++t) cp_async8(dst + (R::rA + t) * 32, in.A + (kk * N * N + t) * L_ + b);
    }
  };
#pragma unroll
  for (int j = 0; j < kBuf - 1; ++j) {
    if (T - j >= 0) fetch(T - j, j);
    cp_async_commit();
  }

  int status = SIPOC_FACTOR_SUCCESS;
  double Ffc[N * N], sdic[N];  // the child's F factor and 1 / sqrt(delta)
  int buf = 0;
  for (int node = T; node >= 0; --node) {
    {
      const int kf = node - (kBuf - 1);
      if (kf >= 0) fetch(kf, buf == 0 ? kBuf - 1 : buf - 1);
      cp_async_commit();
    }
    cp_async_wait_group<kBuf - 1>();
    const double *S = sm_strict + static_cast<size_t>(buf) * R::kRows * 32 + lane;
#define SR(row) S[(row) * 32]
    double V[N * N];
#pragma unroll
    for (int i = 0; i < N * N; ++i) V[i] = SR(R::rQ + i);  // :658

    if (node < T) {
      double A[N * N], B[N * M];
#pragma unroll
      for (int i = 0; i < N * N; ++i) A[i] = SR(R::rA + i);
#pragma unroll
      for (int i = 0; i < N * M; ++i) B[i] = SR(R::rB + i);
      // compute_regularized_W (lqr.cpp:511-529)
      double W[N * N];
#pragma unroll
      for (int i = 0; i < N * N; ++i) W[i] = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) W[i + i * N] = 1.0;
      chol_solve<N, N>(Ffc, W);
#pragma unroll
      for (int i = 0; i < N * N; ++i) W[i] *= -1.0;
#pragma unroll
      for (int i = 0; i < N; ++i) W[i + i * N] += 1.0;
#pragma unroll
      for (int col = 0; col < N; ++col)
#pragma unroll
        for (int row = 0; row < N; ++row) W[row + col * N] *= sdic[row] * sdic[col];
      // H_child = B^T W (M x N)  (:692)
      double H[M * N];
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int a = 0; a < M; ++a) {
          double s = 0.0;
#pragma unroll
          for (int p = 0; p < N; ++p) s += B[p + a * N] * W[p + j * N];
          H[a + j * M] = s;
        }
      // G = R + H_child B  (:693-694), Cholesky (:696-701)
      double Gf[M * M];
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int a = 0; a < M; ++a) {
          double s = SR(R::rR + a + j * M);
#pragma unroll
          for (int p = 0; p < N; ++p) s += H[a + p * M] * B[p + j * N];
          Gf[a + j * M] = s;
        }
      if (!chol_lower<M>(Gf) && status == SIPOC_FACTOR_SUCCESS)
        status = SIPOC_FACTOR_G_FACTORIZATION_FAILURE;
      // F = W A (N x N)  (:703)
      double F[N * N];
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double s = 0.0;
#pragma unroll
          for (int p = 0; p < N; ++p) s += W[i + p * N] * A[p + j * N];
          F[i + j * N] = s;
        }
      // H_parent = M^T + B^T F (M x N)  (:704-705)
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int a = 0; a < M; ++a) {
          double s = SR(R::rM + j + a * N);
#pragma unroll
          for (int p = 0; p < N; ++p) s += B[p + a * N] * F[p + j * N];
          H[a + j * M] = s;
        }
      // K = -G^-1 H_parent  (:707-713)
      double K[M * N];
#pragma unroll
      for (int i = 0; i < M * N; ++i) K[i] = H[i];
      chol_solve<M, N>(Gf, K);
#pragma unroll
      for (int i = 0; i < M * N; ++i) K[i] *= -1.0;
      // V += A^T F  (:715)
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double s = V[i + j * N];
#pragma unroll
          for (int p = 0; p < N; ++p) s += A[p + i * N] * F[p + j * N];
          V[i + j * N] = s;
        }
      // V += K^T H_parent  (:716-719)
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < M; ++a) s += K[a + i * M] * H[a + j * M];
          V[i + j * N] += s;
        }
#pragma unroll
      for (int i = 0; i < N * N; ++i) put(Z::oW(T) + int64_t(node) * N * N + i, W[i]);
#pragma unroll
      for (int i = 0; i < M * N; ++i) put(Z::oK(T) + int64_t(node) * M * N + i, K[i]);
#pragma unroll
      for (int i = 0; i < M * M; ++i) put(Z::oG(T) + int64_t(node) * M * M + i, Gf[i]);
    }

    // compute_delta_sqrt (:475-485), factor_F (:487-509)
    double sd[N];
    bool d_ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double d = SR(R::rD + i);
      d_ok = d_ok && (d > 0.0);
      const double s = sqrt(d);
      sd[i] = s;
      sdic[i] = 1.0 / s;
    }
    if (!d_ok && status == SIPOC_FACTOR_SUCCESS) status = SIPOC_FACTOR_INVALID_DELTA;
#pragma unroll
    for (int col = 0; col < N; ++col) {
#pragma unroll
      for (int row = 0; row < N; ++row) Ffc[row + col * N] = sd[row] * V[row + col * N] * sd[col];
      Ffc[col + col * N] += 1.0;
    }
    if (!chol_lower<N>(Ffc) && status == SIPOC_FACTOR_SUCCESS)
      status = SIPOC_FACTOR_F_FACTORIZATION_FAILURE;
#pragma unroll
    for (int i = 0; i < N * N; ++i) {
      put(Z::oV(T) + int64_t(node) * N * N + i, V[i]);
      put(Z::oF(T) + int64_t(node) * N * N + i, Ffc[i]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      put(Z::oSd(T) + int64_t(node) * N + i, sd[i]);
      put(Z::oSdi(T) + int64_t(node) * N + i, sdic[i]);
    }
#undef SR
    buf = buf + 1 == kBuf ? 0 : buf + 1;
  }
  if (status_out != nullptr && valid) status_out[b] = status;
}

// ===========================================================================
// solve (lqr.cpp:735-871) on a chain: backward affine sweep, root, forward rollout.
// ===========================================================================
template <int N, int M>
struct SolveRows {
  // backward sweep, edge k: W, K, G factor, A, B, q_k, r_k, c_{k+1}, delta_{k+1}
  static constexpr int bW = 0, bK = bW + N * N, bG = bK + M * N, bA = bG + M * M, bB = bA + N * N,
                       bq = bB + N * M, br = bq + N, bc = br + M, bd = bc + N, kBack = bd + N;
  // forward rollout, edge k: K, A, B, c, delta, V, F factor, sd, sdi of node k + 1
  static constexpr int fK = 0, fA = fK + M * N, fB = fA + N * N, fc = fB + N * M, fd = fc + N,
                       fV = fd + N, fF = fV + N * N, fs = fF + N * N, fi = fs + N, kFwd = fi + N;
  static constexpr int kRows = kBack > kFwd ? kBack : kFwd;
  static constexpr int kBytes = kBuf * kRows * 32 * int(sizeof(double));
};

template <int N, int M>
__global__ void __launch_bounds__(32)
strict_solve_thread(LqrIn in, LqrOut out, const double *store, double *scratch, int64_t batch,
                    int64_t ld, int T) {
  using Z = StrictSizes<N, M>;
  using R = SolveRows<N, M>;
  extern __shared__ __align__(16) double sm_strict[];
  const int lane = threadIdx.x;
  const int64_t b_raw = static_cast<int64_t>(blockIdx.x) * 32 + lane;
  const bool valid = b_raw < batch;
  const int64_t b = valid ? b_raw : batch - 1;
  const size_t L_ = static_cast<size_t>(ld);
  const double *st = store + b;
  double *vst = scratch + Z::ov(T) * ld + b;
  double *kst = scratch + Z::ok(T) * ld + b;
  auto ld1 = [&](const double *p, size_t flat) { return __ldcs(p + flat * L_); };

  // ---- backward affine sweep (:738-796) ------------------------------------
  auto fetch_back = [&](int k, int buf) {
    double *dst = sm_strict + static_cast<size_t>(buf) * R::kRows * 32 + lane;
    const size_t kk = static_cast<size_t>(k);
#pragma unroll
    for (int t = 0; t < N * N; ++t) {
      cp_async8(dst + (R::bW + t) * 32, st + (Z::oW(T) + kk * N * N + t) * L_);
      cp_async8(dst + (R::bA + t) * 32, in.A + (kk * N * N + t) * L_ + b);
    }
#pragma unroll
    for (int t = 0; t < M * N; ++t) {
      cp_async8(dst + (R::bK + t) * 32, st + (Z::oK(T) + kk * M * N + t) * L_);
      cp_async8(dst + (R::bB + t) * 32, in.B + (kk * N * M + t) * L_ + b);
    }
#pragma unroll
    for (int t = 0; t < M * M; ++t) cp_async8(dst + (R::bG + t) * 32, st + (Z::oG(T) + kk * M * M + t) * L_);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      cp_async8(dst + (R::bq + i) * 32, in.q + (kk * N + i) * L_ + b);
      cp_async8(dst + (R::bc + i) * 32, in.c + ((kk + 1) * N + i) * L_ + b);
      cp_async8(dst + (R::bd + i) * 32, in.delta + ((kk + 1) * N + i) * L_ + b);
    }
#pragma unroll
    for (int a = 0; a < M; ++a) cp_async8(dst + (R::br + a) * 32, in.r + (kk * M + a) * L_ + b);
  };
#pragma unroll
  for (int j = 0; j < kBuf - 1; ++j) {
    if (T - 1 - j >= 0) fetch_back(T - 1 - j, j);
    cp_async_commit();
  }
  double vc[N];  // v of the child node
#pragma unroll
  for (int i = 0; i < N; ++i) {
    vc[i] = ld1(in.q + b, static_cast<size_t>(T) * N + i);
    if (valid) vst[(static_cast<size_t>(T) * N + i) * L_] = vc[i];
  }
  int buf = 0;
  for (int k = T - 1; k >= 0; --k) {
    {
      const int kf = k - (kBuf - 1);
      if (kf >= 0) fetch_back(kf, buf == 0 ? kBuf - 1 : buf - 1);
      cp_async_commit();
    }
    cp_async_wait_group<kBuf - 1>();
    const double *S = sm_strict + static_cast<size_t>(buf) * R::kRows * 32 + lane;
#define SR(row) S[(row) * 32]
    double f[N], g[N], h[M];
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = SR(R::bd + i) * vc[i] - SR(R::bc + i);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s += SR(R::bW + i + j * N) * f[j];
      g[i] = vc[i] - s;
    }
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) s += SR(R::bB + i + a * N) * g[i];
      h[a] = SR(R::br + a) + s;
    }
    double Gf[M * M], kk[M];
#pragma unroll
    for (int i = 0; i < M * M; ++i) Gf[i] = SR(R::bG + i);
#pragma unroll
    for (int a = 0; a < M; ++a) kk[a] = h[a];
    chol_solve<M, 1>(Gf, kk);
#pragma unroll
    for (int a = 0; a < M; ++a) {
      kk[a] *= -1.0;
      if (valid) kst[(static_cast<size_t>(k) * M + a) * L_] = kk[a];
    }
    double vp[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) s += SR(R::bA + i + j * N) * g[i];
      double w2 = 0.0;
#pragma unroll
      for (int a = 0; a < M; ++a) w2 += SR(R::bK + a + j * M) * h[a];
      double v = SR(R::bq + j);
      v += s;
      v += w2;
      vp[j] = v;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      vc[j] = vp[j];
      if (valid) vst[(static_cast<size_t>(k) * N + j) * L_] = vp[j];
    }
#undef SR
    buf = buf + 1 == kBuf ? 0 : buf + 1;
  }
  cp_async_wait_group<0>();

  // ---- forward rollout (:821-870); its first fetches fly during the root solve -------
  auto fetch_fwd = [&](int k, int bf) {
    double *dst = sm_strict + static_cast<size_t>(bf) * R::kRows * 32 + lane;
    const size_t kk = static_cast<size_t>(k);
#pragma unroll
    for (int t = 0; t < M * N; ++t) {
      cp_async8(dst + (R::fK + t) * 32, st + (Z::oK(T) + kk * M * N + t) * L_);
      cp_async8(dst + (R::fB + t) * 32, in.B + (kk * N * M + t) * L_ + b);
    }
#pragma unroll
    for (int t = 0; t < N * N; ++t) {
      cp_async8(dst + (R::fA + t) * 32, in.A + (kk * N * N + t) * L_ + b);
      cp_async8(dst + (R::fV + t) * 32, st + (Z::oV(T) + (kk + 1) * N * N + t) * L_);
      cp_async8(dst + (R::fF + t) * 32, st + (Z::oF(T) + (kk + 1) * N * N + t) * L_);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      cp_async8(dst + (R::fc + i) * 32, in.c + ((kk + 1) * N + i) * L_ + b);
      cp_async8(dst + (R::fd + i) * 32, in.delta + ((kk + 1) * N + i) * L_ + b);
      cp_async8(dst + (R::fs + i) * 32, st + (Z::oSd(T) + (kk + 1) * N + i) * L_);
      cp_async8(dst + (R::fi + i) * 32, st + (Z::oSdi(T) + (kk + 1) * N + i) * L_);
    }
  };
#pragma unroll
  for (int j = 0; j < kBuf - 1; ++j) {
    if (j < T) fetch_fwd(j, j);
    cp_async_commit();
  }
  // (I + D V)^-1 rhs = D^1/2 F^-1 D^-1/2 rhs  (f_inv_mult, lqr.cpp:531-549)
  auto f_inv_mult = [&](const double (&Ff)[N * N], const double (&sd)[N], const double (&sdi)[N],
                        const double (&rhs)[N], double (&res)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) res[i] = sdi[i] * rhs[i];
    chol_solve<N, 1>(Ff, res);
#pragma unroll
    for (int i = 0; i < N; ++i) res[i] *= sd[i];
  };
  double x[N];
  {  // root (:798-819): vc is v_0
    double Ff[N * N], V0[N * N], sd[N], sdi[N], f[N];
#pragma unroll
    for (int i = 0; i < N * N; ++i) {
      Ff[i] = ld1(st, Z::oF(T) + i);
      V0[i] = ld1(st, Z::oV(T) + i);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      sd[i] = ld1(st, Z::oSd(T) + i);
      sdi[i] = ld1(st, Z::oSdi(T) + i);
      f[i] = ld1(in.delta + b, i) * vc[i] - ld1(in.c + b, i);
    }
    f_inv_mult(Ff, sd, sdi, f, x);
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] *= -1.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s += V0[i + j * N] * x[j];
      if (valid) {
        __stcs(out.x + static_cast<size_t>(i) * L_ + b, x[i]);
        __stcs(out.y + static_cast<size_t>(i) * L_ + b, vc[i] + s);
      }
    }
  }
  buf = 0;
  for (int k = 0; k < T; ++k) {
    {
      const int kf = k + (kBuf - 1);
      if (kf < T) fetch_fwd(kf, buf == 0 ? kBuf - 1 : buf - 1);
      cp_async_commit();
    }
    cp_async_wait_group<kBuf - 1>();
    const double *S = sm_strict + static_cast<size_t>(buf) * R::kRows * 32 + lane;
#define SR(row) S[(row) * 32]
    double u[M], f[N], Ff[N * N], sd[N], sdi[N], vch[N];
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s += SR(R::fK + a + j * M) * x[j];
      u[a] = kst[(static_cast<size_t>(k) * M + a) * L_] + s;
      if (valid) __stcs(out.u + (static_cast<size_t>(k) * M + a) * L_ + b, u[a]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      vch[i] = vst[(static_cast<size_t>(k + 1) * N + i) * L_];
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s += SR(R::fA + i + j * N) * x[j];
      double w2 = 0.0;
#pragma unroll
      for (int a = 0; a < M; ++a) w2 += SR(R::fB + i + a * N) * u[a];
      double fi = SR(R::fc + i) - SR(R::fd + i) * vch[i];
      fi += s;
      fi += w2;
      f[i] = fi;
      sd[i] = SR(R::fs + i);
      sdi[i] = SR(R::fi + i);
    }
#pragma unroll
    for (int i = 0; i < N * N; ++i) Ff[i] = SR(R::fF + i);
    f_inv_mult(Ff, sd, sdi, f, x);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s += SR(R::fV + i + j * N) * x[j];
      if (valid) {
        __stcs(out.x + (static_cast<size_t>(k + 1) * N + i) * L_ + b, x[i]);
        __stcs(out.y + (static_cast<size_t>(k + 1) * N + i) * L_ + b, vch[i] + s);
      }
    }
#undef SR
    buf = buf + 1 == kBuf ? 0 : buf + 1;
  }
}


// ===========================================================================
// The same two routines on a TREE (padded to a uniform shape like the chains): nodes in
// post-order, every node folding in all of its child edges (lqr.cpp:660-720), the affine
// sweep likewise (:746-795), the rollout in pre-order (:827-869).  What a chain carries
// from one stage to the next in registers (the child's F factor, 1 / sqrt(delta), v, x) is
// read back from the thread's own earlier stores here; operands are plain coalesced loads,
// all of an item's loads in flight together (fully unrolled, no dependent addressing).
// ===========================================================================
template <int N, int M>
__global__ void __launch_bounds__(32)
strict_factor_tree(DevTables t, LqrIn in, int *status_out, double *store, int64_t batch,
                   int64_t ld) {
  const int T = t.E;
  using Z = StrictSizes<N, M>;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * 32 + threadIdx.x;
  if (b >= batch) return;
  const size_t L_ = static_cast<size_t>(ld);
  double *st = store + b;
  auto ld1 = [&](const double *p, size_t flat) { return __ldcs(p + flat * L_ + b); };
  auto lds = [&](size_t flat) { return st[flat * L_]; };  // own earlier stores
  auto put = [&](size_t flat, double v) { st[flat * L_] = v; };

  int status = SIPOC_FACTOR_SUCCESS;
  for (int order = 0; order < t.N; ++order) {
    const int node = t.postorder[order];
    double V[N * N];
#pragma unroll
    for (int i = 0; i < N * N; ++i) V[i] = ld1(in.Q, static_cast<size_t>(node) * N * N + i);  // :658

    for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {
      const size_t e = t.child_edges[ci], child = t.children[e];
      double A[N * N], B[N * M], Ffc[N * N], sdic[N];
#pragma unroll
      for (int i = 0; i < N * N; ++i) {
        A[i] = ld1(in.A, e * N * N + i);
        Ffc[i] = lds(Z::oF(T) + child * N * N + i);
      }
#pragma unroll
      for (int i = 0; i < N * M; ++i) B[i] = ld1(in.B, e * N * M + i);
#pragma unroll
      for (int i = 0; i < N; ++i) sdic[i] = lds(Z::oSdi(T) + child * N + i);
      // compute_regularized_W (lqr.cpp:511-529)
      double W[N * N];
#pragma unroll
      for (int i = 0; i < N * N; ++i) W[i] = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) W[i + i * N] = 1.0;
      chol_solve<N, N>(Ffc, W);
#pragma unroll
      for (int i = 0; i < N * N; ++i) W[i] *= -1.0;
#pragma unroll
      for (int i = 0; i < N; ++i) W[i + i * N] += 1.0;
#pragma unroll
      for (int col = 0; col < N; ++col)
#pragma unroll
        for (int row = 0; row < N; ++row) W[row + col * N] *= sdic[row] * sdic[col];
      double H[M * N];
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int a = 0; a < M; ++a) {
          double s = 0.0;
#pragma unroll
          for (int p = 0; p < N; ++p) s += B[p + a * N] * W[p + j * N];
          H[a + j * M] = s;
        }
      double Gf[M * M];
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int a = 0; a < M; ++a) {
          double s = ld1(in.R, e * M * M + a + j * M);
#pragma unroll
          for (int p = 0; p < N; ++p) s += H[a + p * M] * B[p + j * N];
          Gf[a + j * M] = s;
        }
      if (!chol_lower<M>(Gf) && status == SIPOC_FACTOR_SUCCESS)
        status = SIPOC_FACTOR_G_FACTORIZATION_FAILURE;
      double F[N * N];
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double s = 0.0;
#pragma unroll
          for (int p = 0; p < N; ++p) s += W[i + p * N] * A[p + j * N];
          F[i + j * N] = s;
        }
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int a = 0; a < M; ++a) {
          double s = ld1(in.M, e * N * M + j + a * N);
#pragma unroll
          for (int p = 0; p < N; ++p) s += B[p + a * N] * F[p + j * N];
          H[a + j * M] = s;
        }
      double K[M * N];
#pragma unroll
      for (int i = 0; i < M * N; ++i) K[i] = H[i];
      chol_solve<M, N>(Gf, K);
#pragma unroll
      for (int i = 0; i < M * N; ++i) K[i] *= -1.0;
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double s = V[i + j * N];
#pragma unroll
          for (int p = 0; p < N; ++p) s += A[p + i * N] * F[p + j * N];
          V[i + j * N] = s;
        }
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < M; ++a) s += K[a + i * M] * H[a + j * M];
          V[i + j * N] += s;
        }
#pragma unroll
      for (int i = 0; i < N * N; ++i) put(Z::oW(T) + e * N * N + i, W[i]);
#pragma unroll
      for (int i = 0; i < M * N; ++i) put(Z::oK(T) + e * M * N + i, K[i]);
#pragma unroll
      for (int i = 0; i < M * M; ++i) put(Z::oG(T) + e * M * M + i, Gf[i]);
    }

    double sd[N], sdi[N], Ff[N * N];
    bool d_ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double d = ld1(in.delta, static_cast<size_t>(node) * N + i);
      d_ok = d_ok && (d > 0.0);
      const double s = sqrt(d);
      sd[i] = s;
      sdi[i] = 1.0 / s;
    }
    if (!d_ok && status == SIPOC_FACTOR_SUCCESS) status = SIPOC_FACTOR_INVALID_DELTA;
#pragma unroll
    for (int col = 0; col < N; ++col) {
#pragma unroll
      for (int row = 0; row < N; ++row) Ff[row + col * N] = sd[row] * V[row + col * N] * sd[col];
      Ff[col + col * N] += 1.0;
    }
    if (!chol_lower<N>(Ff) && status == SIPOC_FACTOR_SUCCESS)
      status = SIPOC_FACTOR_F_FACTORIZATION_FAILURE;
#pragma unroll
    for (int i = 0; i < N * N; ++i) {
      put(Z::oV(T) + static_cast<size_t>(node) * N * N + i, V[i]);
      put(Z::oF(T) + static_cast<size_t>(node) * N * N + i, Ff[i]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      put(Z::oSd(T) + static_cast<size_t>(node) * N + i, sd[i]);
      put(Z::oSdi(T) + static_cast<size_t>(node) * N + i, sdi[i]);
    }
  }
  if (status_out != nullptr) status_out[b] = status;
}

template <int N, int M>
__global__ void __launch_bounds__(32)
strict_solve_tree(DevTables t, LqrIn in, LqrOut out, const double *store, double *scratch,
                  int64_t batch, int64_t ld) {
  const int T = t.E;
  using Z = StrictSizes<N, M>;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * 32 + threadIdx.x;
  if (b >= batch) return;
  const size_t L_ = static_cast<size_t>(ld);
  const double *st = store + b;
  double *vst = scratch + Z::ov(T) * ld + b;
  double *kst = scratch + Z::ok(T) * ld + b;
  auto ld1 = [&](const double *p, size_t flat) { return __ldcs(p + flat * L_ + b); };
  auto lds = [&](size_t flat) { return __ldcs(st + flat * L_); };

  // backward affine sweep (:738-796)
  for (int order = 0; order < t.N; ++order) {
    const size_t node = t.postorder[order];
    double v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = ld1(in.q, node * N + i);
    for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {
      const size_t e = t.child_edges[ci], child = t.children[e];
      double vc[N], f[N], g[N], h[M];
#pragma unroll
      for (int i = 0; i < N; ++i) vc[i] = vst[(child * N + i) * L_];
#pragma unroll
      for (int i = 0; i < N; ++i)
        f[i] = ld1(in.delta, child * N + i) * vc[i] - ld1(in.c, child * N + i);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) s += lds(Z::oW(T) + e * N * N + i + j * N) * f[j];
        g[i] = vc[i] - s;
      }
#pragma unroll
      for (int a = 0; a < M; ++a) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) s += ld1(in.B, e * N * M + i + a * N) * g[i];
        h[a] = ld1(in.r, e * M + a) + s;
      }
      double Gf[M * M], kk[M];
#pragma unroll
      for (int i = 0; i < M * M; ++i) Gf[i] = lds(Z::oG(T) + e * M * M + i);
#pragma unroll
      for (int a = 0; a < M; ++a) kk[a] = h[a];
      chol_solve<M, 1>(Gf, kk);
#pragma unroll
      for (int a = 0; a < M; ++a) {
        kk[a] *= -1.0;
        kst[(e * M + a) * L_] = kk[a];
      }
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) s += ld1(in.A, e * N * N + i + j * N) * g[i];
        double w2 = 0.0;
#pragma unroll
        for (int a = 0; a < M; ++a) w2 += lds(Z::oK(T) + e * M * N + a + j * M) * h[a];
        v[j] += s;
        v[j] += w2;
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) vst[(node * N + i) * L_] = v[i];
  }

  auto f_inv_mult = [&](const double (&Ff)[N * N], const double (&sd)[N], const double (&sdi)[N],
                        const double (&rhs)[N], double (&res)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) res[i] = sdi[i] * rhs[i];
    chol_solve<N, 1>(Ff, res);
#pragma unroll
    for (int i = 0; i < N; ++i) res[i] *= sd[i];
  };
  // x, y of `node` from the right-hand side f (negate = the root's sign, :812-819)
  auto close_node = [&](size_t node, const double (&f)[N], const double (&vn)[N], bool negate) {
    double Ff[N * N], sd[N], sdi[N], x[N];
#pragma unroll
    for (int i = 0; i < N * N; ++i) Ff[i] = lds(Z::oF(T) + node * N * N + i);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      sd[i] = lds(Z::oSd(T) + node * N + i);
      sdi[i] = lds(Z::oSdi(T) + node * N + i);
    }
    f_inv_mult(Ff, sd, sdi, f, x);
    if (negate) {
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] *= -1.0;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s += lds(Z::oV(T) + node * N * N + i + j * N) * x[j];
      out.x[(node * N + i) * L_ + b] = x[i];
      out.y[(node * N + i) * L_ + b] = vn[i] + s;
    }
  };
  {  // root
    const size_t root = t.preorder[0];
    double f[N], vr[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      vr[i] = vst[(root * N + i) * L_];
      f[i] = ld1(in.delta, root * N + i) * vr[i] - ld1(in.c, root * N + i);
    }
    close_node(root, f, vr, true);
  }
  // forward rollout (:821-870)
  for (int order = 0; order < t.N; ++order) {
    const size_t node = t.preorder[order];
    double x[N];
    bool have_x = false;
    for (int ci = t.child_offsets[node]; ci < t.child_offsets[node + 1]; ++ci) {
      const size_t e = t.child_edges[ci], child = t.children[e];
      if (!have_x) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = out.x[(node * N + i) * L_ + b];
        have_x = true;
      }
      double u[M], f[N], vch[N];
#pragma unroll
      for (int a = 0; a < M; ++a) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) s += lds(Z::oK(T) + e * M * N + a + j * M) * x[j];
        u[a] = kst[(e * M + a) * L_] + s;
        out.u[(e * M + a) * L_ + b] = u[a];
      }
#pragma unroll
      for (int i = 0; i < N; ++i) {
        vch[i] = vst[(child * N + i) * L_];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) s += ld1(in.A, e * N * N + i + j * N) * x[j];
        double w2 = 0.0;
#pragma unroll
        for (int a = 0; a < M; ++a) w2 += ld1(in.B, e * N * M + i + a * N) * u[a];
        double fi = ld1(in.c, child * N + i) - ld1(in.delta, child * N + i) * vch[i];
        fi += s;
        fi += w2;
        f[i] = fi;
      }
      close_node(child, f, vch, false);
    }
  }
}

template <int N, int M>
struct StrictPlan {
  static int64_t store_elems(int T) { return StrictSizes<N, M>::store(T); }
  static int64_t scratch_elems(int T) { return StrictSizes<N, M>::scratch(T); }
  static int factor(const FastArgs &a, cudaStream_t s) {
    auto kern = strict_factor_thread<N, M>;
    constexpr int bytes = FactorRows<N, M>::kBytes;
    if (bytes > 48 * 1024)
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    ProfScope ps(a.prof, "strict_factor_thread", s);
    kern<<<static_cast<unsigned>((a.batch + 31) / 32), 32, bytes, s>>>(a.in, a.status, a.store,
                                                                       a.batch, a.ld, a.num_edges);
    return 1;
  }
  static int solve(const FastArgs &a, cudaStream_t s) {
    auto kern = strict_solve_thread<N, M>;
    constexpr int bytes = SolveRows<N, M>::kBytes;
    if (bytes > 48 * 1024)
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    ProfScope ps(a.prof, "strict_solve_thread", s);
    kern<<<static_cast<unsigned>((a.batch + 31) / 32), 32, bytes, s>>>(
        a.in, a.out, a.store, a.scratch, a.batch, a.ld, a.num_edges);
    return 1;
  }
  static int factor_solve(const FastArgs &a, cudaStream_t s) { return factor(a, s) + solve(a, s); }
  // trees (FastArgs::tables carries the topology)
  static int factor_tree(const FastArgs &a, cudaStream_t s) {
    ProfScope ps(a.prof, "strict_factor_tree", s);
    strict_factor_tree<N, M><<<static_cast<unsigned>((a.batch + 31) / 32), 32, 0, s>>>(
        *a.tables, a.in, a.status, a.store, a.batch, a.ld);
    return 1;
  }
  static int solve_tree(const FastArgs &a, cudaStream_t s) {
    ProfScope ps(a.prof, "strict_solve_tree", s);
    strict_solve_tree<N, M><<<static_cast<unsigned>((a.batch + 31) / 32), 32, 0, s>>>(
        *a.tables, a.in, a.out, a.store, a.scratch, a.batch, a.ld);
    return 1;
  }
  static int factor_solve_tree(const FastArgs &a, cudaStream_t s) {
    return factor_tree(a, s) + solve_tree(a, s);
  }
};

template <int N, int M>
const FastPlan *make_strict_plan(const char *name) {
  using P = StrictPlan<N, M>;
  static const FastPlan plan{name,         N,           M,         &P::store_elems,
                             &P::scratch_elems, &P::factor, &P::solve, &P::factor_solve,
                             nullptr, false};
  return &plan;
}

template <int N, int M>
const FastPlan *make_strict_tree_plan(const char *name) {
  using P = StrictPlan<N, M>;
  static const FastPlan plan{name,         N,           M,         &P::store_elems,
                             &P::scratch_elems, &P::factor_tree, &P::solve_tree,
                             &P::factor_solve_tree, nullptr, false};
  return &plan;
}

}  // namespace

// Smallest instantiated reference-order shape that holds (n, m); nullptr when none does.
const FastPlan *select_strict_plan(int n, int m) {
  if (n <= 2 && m <= 1) return make_strict_plan<2, 1>("strict_thread_n2_m1");
  if (n <= 3 && m <= 2) return make_strict_plan<3, 2>("strict_thread_n3_m2");
  if (n <= 4 && m <= 2) return make_strict_plan<4, 2>("strict_thread_n4_m2");
  if (n <= 4 && m <= 4) return make_strict_plan<4, 4>("strict_thread_n4_m4");
  if (n <= 5 && m <= 3) return make_strict_plan<5, 3>("strict_thread_n5_m3");
  if (n <= 6 && m <= 3) return make_strict_plan<6, 3>("strict_thread_n6_m3");
  return nullptr;
}

// The same shapes on trees.
const FastPlan *select_strict_tree_plan(int n, int m) {
  if (n <= 2 && m <= 1) return make_strict_tree_plan<2, 1>("strict_tree_n2_m1");
  if (n <= 3 && m <= 2) return make_strict_tree_plan<3, 2>("strict_tree_n3_m2");
  if (n <= 4 && m <= 2) return make_strict_tree_plan<4, 2>("strict_tree_n4_m2");
  if (n <= 4 && m <= 4) return make_strict_tree_plan<4, 4>("strict_tree_n4_m4");
  if (n <= 5 && m <= 3) return make_strict_tree_plan<5, 3>("strict_tree_n5_m3");
  return nullptr;
}

}  // namespace sipoc
