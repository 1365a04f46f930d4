// Optional FP32 mode (north_star: "an optional FP32 mode is reported separately with its own
// stated tolerance"): regularized-LQR factor + solve on uniform chains with tiny state
// dimension (n = 4, m = 1..4 -- the cartpole row of the reference benchmark grid,
// lqr_benchmark.cpp:537-545), every array in single precision in the engine layout
// [flat][ld], one thread per problem with every matrix of the stage in registers.
//
// Same recursion as the FP64 thread-per-problem kernels (reference
// lqr.cpp:645-731 factor, :738-870 solve), written on a scalar type T so that the T = double
// instantiation can be held to the FP64 tolerance against the oracle (it validates the
// algebra) while T = float is held to the FP32 one:
//   node:  F = I + D^1/2 V D^1/2 = L L',  P = F^-1 = L^-T L^-1 (kept),  W = D^-1/2 (I - P) D^-1/2
//   edge:  S = W [B | A],  Psi = [R M'; M Q] + [B | A]' S,  G = Psi_uu = L_G L_G',
//          Lam = Psi_xu L_G^-T,  K = -L_G^-T Lam' (kept),  V = Psi_xx - Lam Lam'
//          f = delta' o v' - c',  g = v' - W f,  h = r + B' g,  w = q + A' g,
//          t = L_G^-1 h,  k = -L_G^-T t (kept),  v = w - Lam t (kept)
//   root:  x = -(I + D V)^-1 (delta o v - c),  y = v + V x
//   roll:  u = k + K x,  f = c' - delta' o v' + A x + B u,  s = D'^-1/2 f,  t = P' s,
//          x' = D'^1/2 t  (the reference's F-solve form, lqr.cpp:531-549),  y' = v' + D'^-1/2 (s - t)
// Failure order as the reference's (lqr.cpp:696-700, 722-727): G of an edge, then delta of
// its parent node, then F; the first failure in post-order is reported per problem.
//
// Why FP32 pays here: at 16 384 problems x 100 stages (under one warp per scheduler) the step is
// bound by each thread's own stalls per stage; registers hold three to four stages of float
// operands in flight where they hold one of doubles, and the bytes halve.
#include "riccati_f32.cuh"

namespace sipoc {
namespace {

__host__ __device__ constexpr int tri(int n) { return n * (n + 1) / 2; }
// Column-major packed lower index, i >= j.
__host__ __device__ constexpr int pk(int i, int j, int n) { return j * n - j * (j - 1) / 2 + (i - j); }

template <class T>
__device__ __forceinline__ T rsq(T x);
template <>
__device__ __forceinline__ float rsq<float>(float x) { return rsqrtf(x); }
template <>
__device__ __forceinline__ double rsq<double>(double x) { return rsqrt(x); }

// Per-problem element counts of the kept factorization (P per node, K per edge) and of the
// affine spill (v per node, k per edge).
template <int N, int M>
struct F32Sizes {
  static __host__ __device__ constexpr int64_t oP(int) { return 0; }
  static __host__ __device__ constexpr int64_t oK(int T) { return int64_t(T + 1) * tri(N); }
  static __host__ __device__ constexpr int64_t ov(int T) { return oK(T) + int64_t(T) * N * M; }
  static __host__ __device__ constexpr int64_t ok(int T) { return ov(T) + int64_t(T + 1) * N; }
  static __host__ __device__ constexpr int64_t total(int T) { return ok(T) + int64_t(T) * M; }
};

template <class T, int N, int M>
struct StageInputs {
  T A[N * N], B[N * M], Q[tri(N)], Mx[N * M], R[tri(M)], q[N], r[M], c[N], d[N];
};

// Edge k's matrices and vectors (Q, q, delta of node k -- what the stage's node step reads
// last; c of node k + 1; delta of node k + 1 is carried over from the previous stage).
template <class T, int N, int M>
__device__ __forceinline__ void fetch_stage(StageInputs<T, N, M> &s, const LqrInT<T> &in, int k,
                                            size_t L, int64_t b) {
  // 32-bit element indices (the launcher checks that every array has fewer than 2^31 elements):
  // one IMAD for the index and one IMAD.WIDE for the address per load
  const unsigned L32 = static_cast<unsigned>(L), b32 = static_cast<unsigned>(b);
  auto G = [&](const T *p, int e) { return __ldcs(p + (static_cast<unsigned>(e) * L32 + b32)); };
#pragma unroll
  for (int t = 0; t < N * N; ++t) s.A[t] = G(in.A, k * N * N + t);
#pragma unroll
  for (int t = 0; t < N * M; ++t) {
    s.B[t] = G(in.B, k * N * M + t);
    s.Mx[t] = G(in.M, k * N * M + t);
  }
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = j; i < N; ++i) s.Q[pk(i, j, N)] = G(in.Q, (k * N + j) * N + i);
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int i = j; i < M; ++i) s.R[pk(i, j, M)] = G(in.R, (k * M + j) * M + i);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s.q[i] = G(in.q, k * N + i);
    s.c[i] = G(in.c, (k + 1) * N + i);
    s.d[i] = G(in.delta, k * N + i);
  }
#pragma unroll
  for (int a = 0; a < M; ++a) s.r[a] = G(in.r, k * M + a);
}

// In-place lower Cholesky of a packed n x n block with the reciprocal diagonal in dinv
// (column j of L is V[pk(i, j)] for i > j; the diagonal itself is 1 / dinv[j]).
template <class T, int N>
__device__ __forceinline__ bool chol_packed(T (&V)[tri(N)], T (&dinv)[N]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    T x = V[pk(j, j, N)];
#pragma unroll
    for (int p = 0; p < j; ++p) x -= V[pk(j, p, N)] * V[pk(j, p, N)];
    ok = ok && (x > T(0));
    const T d = rsq<T>(x);
    dinv[j] = d;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      T t = V[pk(i, j, N)];
#pragma unroll
      for (int p = 0; p < j; ++p) t -= V[pk(i, p, N)] * V[pk(j, p, N)];
      V[pk(i, j, N)] = t * d;
    }
  }
  return ok;
}

// s = D^-1/2 z, t = P s:  fz = D^1/2 t = (I + D V)^-1 z,  wz = D^-1/2 (s - t) = W z.
template <class T, int N>
__device__ __forceinline__ void apply_node(const T (&P)[tri(N)], const T (&d)[N], const T (&z)[N],
                                           T (&fz)[N], T (&wz)[N]) {
  T s[N], t[N], sdi[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    sdi[i] = rsq<T>(d[i]);
    s[i] = sdi[i] * z[i];
    t[i] = T(0);
  }
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = j; i < N; ++i) {
      t[i] += P[pk(i, j, N)] * s[j];
      if (i != j) t[j] += P[pk(i, j, N)] * s[i];
    }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    fz[i] = d[i] * sdi[i] * t[i];
    wz[i] = sdi[i] * (s[i] - t[i]);
  }
}

template <class T, int N, int M>
__global__ void __launch_bounds__(64)
lqr_thread_backward(LqrInT<T> in, int *status_out, T *store, int64_t batch, int64_t ld, int Tn) {
  using Z = F32Sizes<N, M>;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const size_t L = static_cast<size_t>(ld);
  T *Pst = store + Z::oP(Tn) * ld + b, *Kst = store + Z::oK(Tn) * ld + b;
  T *vst = store + Z::ov(Tn) * ld + b, *kst = store + Z::ok(Tn) * ld + b;
  const unsigned L32 = static_cast<unsigned>(L);
  auto G = [&](const T *p, int e) {
    return __ldcs(p + (static_cast<unsigned>(e) * L32 + static_cast<unsigned>(b)));
  };
  int status = SIPOC_FACTOR_SUCCESS;
  T W[tri(N)], v[N], dl[N];

  // (V, vv, delta) of node k -> W, v, dl; stores P_k and v_k.
  auto process_node = [&](int k, T (&V)[tri(N)], const T (&vv)[N], const T (&dk)[N]) {
    T sd[N], sdi[N], dinv[N];
    bool d_ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      d_ok = d_ok && (dk[i] > T(0));
      dl[i] = dk[i];
      sdi[i] = rsq<T>(dk[i]);
      sd[i] = dk[i] * sdi[i];
    }
    if (!d_ok && status == SIPOC_FACTOR_SUCCESS) status = SIPOC_FACTOR_INVALID_DELTA;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int i = j; i < N; ++i)
        V[pk(i, j, N)] = sd[i] * V[pk(i, j, N)] * sd[j] + (i == j ? T(1) : T(0));
    const bool f_ok = chol_packed<T, N>(V, dinv);
    if (!f_ok && status == SIPOC_FACTOR_SUCCESS) status = SIPOC_FACTOR_F_FACTORIZATION_FAILURE;
    T Li[tri(N)];  // L^-1, lower
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int i = j; i < N; ++i) {
        T t = (i == j) ? T(1) : T(0);
#pragma unroll
        for (int p = j; p < i; ++p) t -= V[pk(i, p, N)] * Li[pk(p, j, N)];
        Li[pk(i, j, N)] = t * dinv[i];
      }
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int i = j; i < N; ++i) {
        T fin = T(0);
#pragma unroll
        for (int p = i; p < N; ++p) fin += Li[pk(p, i, N)] * Li[pk(p, j, N)];
        W[pk(i, j, N)] = sdi[i] * ((i == j ? T(1) : T(0)) - fin) * sdi[j];
        __stcs(Pst + static_cast<unsigned>(k * tri(N) + pk(i, j, N)) * L32, fin);
      }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      v[i] = vv[i];
      __stcs(vst + static_cast<unsigned>(k * N + i) * L32, vv[i]);
    }
  };

  // kBufs rotating register buffers (float: 3, double: 2 -- what the register file holds):
  // edge k is computed from its buffer while the edges behind it are in flight, and the buffer
  // is refilled with edge k - kBufs as soon as the stage is done with it.  One stage of
  // compute (~0.5 us) is shorter than the memory latency under load, so a single stage of
  // look-ahead leaves every stage waiting on its loads.
  constexpr int kBufs = sizeof(T) == 4 ? 3 : 2;
  StageInputs<T, N, M> buf[kBufs];
#pragma unroll
  for (int d = 0; d < kBufs; ++d)
    if (Tn - 1 - d >= 0) fetch_stage<T, N, M>(buf[d], in, Tn - 1 - d, L, b);
  {  // terminal node: V = Q_T, v = q_T  (lqr.cpp:658, 744)
    T V[tri(N)], vv[N], dk[N];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int i = j; i < N; ++i) V[pk(i, j, N)] = G(in.Q, (Tn * N + j) * N + i);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      vv[i] = G(in.q, Tn * N + i);
      dk[i] = G(in.delta, Tn * N + i);
    }
    process_node(Tn, V, vv, dk);
  }
  auto stage = [&](int k, StageInputs<T, N, M> &cur) {
    // S = W' [B | A]
    T SB[N * M], SA[N * N];
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T s = T(0);
#pragma unroll
        for (int p = 0; p < N; ++p) s += W[i >= p ? pk(i, p, N) : pk(p, i, N)] * cur.B[p + c * N];
        SB[i + c * N] = s;
      }
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T s = T(0);
#pragma unroll
        for (int p = 0; p < N; ++p) s += W[i >= p ? pk(i, p, N) : pk(p, i, N)] * cur.A[p + c * N];
        SA[i + c * N] = s;
      }
    // Psi blocks (lower): uu = R + B' S_B, xu = M + A' S_B, xx = Q + A' S_A
    T Puu[tri(M)], Pxu[N * M], V[tri(N)];
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int a = c; a < M; ++a) {
        T s = cur.R[pk(a, c, M)];
#pragma unroll
        for (int p = 0; p < N; ++p) s += cur.B[p + a * N] * SB[p + c * N];
        Puu[pk(a, c, M)] = s;
      }
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T s = cur.Mx[i + c * N];
#pragma unroll
        for (int p = 0; p < N; ++p) s += cur.A[p + i * N] * SB[p + c * N];
        Pxu[i + c * N] = s;
      }
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
      for (int i = c; i < N; ++i) {
        T s = cur.Q[pk(i, c, N)];
#pragma unroll
        for (int p = 0; p < N; ++p) s += cur.A[p + i * N] * SA[p + c * N];
        V[pk(i, c, N)] = s;
      }
    // G = L_G L_G'
    T gdinv[M];
    const bool g_ok = chol_packed<T, M>(Puu, gdinv);
    if (!g_ok && status == SIPOC_FACTOR_SUCCESS) status = SIPOC_FACTOR_G_FACTORIZATION_FAILURE;
    // Lam = Psi_xu L_G^-T : row i solves L_G lam_i = pxu_i
    T Lam[N * M];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int a = 0; a < M; ++a) {
        T t = Pxu[i + a * N];
#pragma unroll
        for (int p = 0; p < a; ++p) t -= Puu[pk(a, p, M)] * Lam[i + p * N];
        Lam[i + a * N] = t * gdinv[a];
      }
    // K = -L_G^-T Lam'  (m x n), column i of K from row i of Lam
#pragma unroll
    for (int i = 0; i < N; ++i) {
      T kap[M];
#pragma unroll
      for (int a = M - 1; a >= 0; --a) {
        T t = Lam[i + a * N];
#pragma unroll
        for (int p = a + 1; p < M; ++p) t -= Puu[pk(p, a, M)] * kap[p];
        kap[a] = t * gdinv[a];
      }
#pragma unroll
      for (int a = 0; a < M; ++a)
        __stcs(Kst + static_cast<unsigned>(k * N * M + a + i * M) * L32, -kap[a]);
    }
    // V = Psi_xx - Lam Lam'
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
      for (int i = c; i < N; ++i) {
        T s = V[pk(i, c, N)];
#pragma unroll
        for (int a = 0; a < M; ++a) s -= Lam[i + a * N] * Lam[c + a * N];
        V[pk(i, c, N)] = s;
      }
    // affine part
    T f[N], g[N], h[M], w[N], tt[M], kk[M], vv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = dl[i] * v[i] - cur.c[i];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      T s = T(0);
#pragma unroll
      for (int p = 0; p < N; ++p) s += W[i >= p ? pk(i, p, N) : pk(p, i, N)] * f[p];
      g[i] = v[i] - s;
    }
#pragma unroll
    for (int a = 0; a < M; ++a) {
      T s = cur.r[a];
#pragma unroll
      for (int p = 0; p < N; ++p) s += cur.B[p + a * N] * g[p];
      h[a] = s;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      T s = cur.q[i];
#pragma unroll
      for (int p = 0; p < N; ++p) s += cur.A[p + i * N] * g[p];
      w[i] = s;
    }
#pragma unroll
    for (int a = 0; a < M; ++a) {
      T t = h[a];
#pragma unroll
      for (int p = 0; p < a; ++p) t -= Puu[pk(a, p, M)] * tt[p];
      tt[a] = t * gdinv[a];
    }
#pragma unroll
    for (int a = M - 1; a >= 0; --a) {
      T t = tt[a];
#pragma unroll
      for (int p = a + 1; p < M; ++p) t -= Puu[pk(p, a, M)] * kk[p];
      kk[a] = t * gdinv[a];
    }
#pragma unroll
    for (int a = 0; a < M; ++a) __stcs(kst + static_cast<unsigned>(k * M + a) * L32, -kk[a]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      T s = w[i];
#pragma unroll
      for (int a = 0; a < M; ++a) s -= Lam[i + a * N] * tt[a];
      vv[i] = s;
    }
    // delta of node k came with this stage's own operands (a value out of another buffer's
    // fetch made the stage wait on that buffer's scoreboard, i.e. on loads issued later)
    T dk[N];
#pragma unroll
    for (int i = 0; i < N; ++i) dk[i] = cur.d[i];
    if (k - kBufs >= 0) fetch_stage<T, N, M>(cur, in, k - kBufs, L, b);  // the buffer is free
    process_node(k, V, vv, dk);
  };
  for (int k0 = Tn - 1; k0 >= 0; k0 -= kBufs) {
#pragma unroll
    for (int slot = 0; slot < kBufs; ++slot) {
      if (k0 - slot < 0) break;
      stage(k0 - slot, buf[slot]);
    }
  }
  if (status_out != nullptr) status_out[b] = status;
}

template <class T, int N, int M>
struct RollInputs {
  T A[N * N], B[N * M], K[N * M], P[tri(N)], c[N], d[N], v[N], k[M];
};

template <class T, int N, int M>
__device__ __forceinline__ void fetch_roll(RollInputs<T, N, M> &s, const LqrInT<T> &in,
                                           const T *store, int k, int Tn, size_t L, int64_t b) {
  using Z = F32Sizes<N, M>;
  const unsigned L32 = static_cast<unsigned>(L), b32 = static_cast<unsigned>(b);
  auto G = [&](const T *p, size_t e) { return __ldcs(p + (static_cast<unsigned>(e) * L32 + b32)); };
#pragma unroll
  for (int t = 0; t < N * N; ++t) s.A[t] = G(in.A, static_cast<size_t>(k) * N * N + t);
#pragma unroll
  for (int t = 0; t < N * M; ++t) {
    s.B[t] = G(in.B, static_cast<size_t>(k) * N * M + t);
    s.K[t] = G(store, Z::oK(Tn) + static_cast<size_t>(k) * N * M + t);
  }
#pragma unroll
  for (int t = 0; t < tri(N); ++t) s.P[t] = G(store, Z::oP(Tn) + static_cast<size_t>(k + 1) * tri(N) + t);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s.c[i] = G(in.c, static_cast<size_t>(k + 1) * N + i);
    s.d[i] = G(in.delta, static_cast<size_t>(k + 1) * N + i);
    s.v[i] = G(store, Z::ov(Tn) + static_cast<size_t>(k + 1) * N + i);
  }
#pragma unroll
  for (int a = 0; a < M; ++a) s.k[a] = G(store, Z::ok(Tn) + static_cast<size_t>(k) * M + a);
}

template <class T, int N, int M>
__global__ void __launch_bounds__(64)
lqr_thread_rollout(LqrInT<T> in, LqrOutT<T> out, const T *store, int64_t batch, int64_t ld,
                   int Tn) {
  using Z = F32Sizes<N, M>;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const size_t L = static_cast<size_t>(ld);
  const unsigned L32 = static_cast<unsigned>(L), b32 = static_cast<unsigned>(b);
  auto G = [&](const T *p, size_t e) { return __ldcs(p + (static_cast<unsigned>(e) * L32 + b32)); };
  auto S = [&](T *p, int e, T val) { __stcs(p + (static_cast<unsigned>(e) * L32 + b32), val); };
  // kDepth rotating register buffers: a stage of the rollout is ~100 FMAs, far shorter than
  // the memory latency, so the kernel runs at latency / kDepth per stage until the FMAs bound it.
  constexpr int kDepth = sizeof(T) == 4 ? 4 : 2;
  RollInputs<T, N, M> buf[kDepth];
#pragma unroll
  for (int d = 0; d < kDepth; ++d)
    if (d < Tn) fetch_roll<T, N, M>(buf[d], in, store, d, Tn, L, b);
  T x[N];
  {  // root (lqr.cpp:798-819): x = -(I + D V)^-1 (delta o v - c),  y = v + V x = v - W (delta o v - c)
    T P0[tri(N)], d0[N], z[N], fz[N], wz[N], v0[N];
#pragma unroll
    for (int t = 0; t < tri(N); ++t) P0[t] = G(store, Z::oP(Tn) + t);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      d0[i] = G(in.delta, i);
      v0[i] = G(store, Z::ov(Tn) + i);
      z[i] = G(in.c, i) - d0[i] * v0[i];
    }
    apply_node<T, N>(P0, d0, z, fz, wz);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      x[i] = fz[i];
      S(out.x, i, fz[i]);
      S(out.y, i, v0[i] + wz[i]);
    }
  }
  for (int k0 = 0; k0 < Tn; k0 += kDepth) {
#pragma unroll
    for (int slot = 0; slot < kDepth; ++slot) {
    const int k = k0 + slot;
    if (k >= Tn) break;
    RollInputs<T, N, M> &cur = buf[slot];
    T u[M], f[N], fz[N], wz[N];
#pragma unroll
    for (int a = 0; a < M; ++a) {
      T s = cur.k[a];
#pragma unroll
      for (int i = 0; i < N; ++i) s += cur.K[a + i * M] * x[i];
      u[a] = s;
      S(out.u, k * M + a, s);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      T s = cur.c[i] - cur.d[i] * cur.v[i];
#pragma unroll
      for (int p = 0; p < N; ++p) s += cur.A[i + p * N] * x[p];
#pragma unroll
      for (int a = 0; a < M; ++a) s += cur.B[i + a * N] * u[a];
      f[i] = s;
    }
    apply_node<T, N>(cur.P, cur.d, f, fz, wz);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      x[i] = fz[i];
      S(out.x, (k + 1) * N + i, fz[i]);
      S(out.y, (k + 1) * N + i, cur.v[i] + wz[i]);
    }
    if (k + kDepth < Tn) fetch_roll<T, N, M>(cur, in, store, k + kDepth, Tn, L, b);
    }
  }
}

template <class T, int N, int M>
int launch_shape(const LqrInT<T> &in, const LqrOutT<T> &out, int *status, T *store, int64_t batch,
                 int64_t ld, int Tn, Profiler *prof, cudaStream_t s) {
  const unsigned grid = static_cast<unsigned>((batch + 63) / 64);
  {
    ProfScope ps(prof, sizeof(T) == 4 ? "lqr_thread_backward_f32" : "lqr_thread_backward_f64", s);
    lqr_thread_backward<T, N, M><<<grid, 64, 0, s>>>(in, status, store, batch, ld, Tn);
  }
  {
    ProfScope ps(prof, sizeof(T) == 4 ? "lqr_thread_rollout_f32" : "lqr_thread_rollout_f64", s);
    lqr_thread_rollout<T, N, M><<<grid, 64, 0, s>>>(in, out, store, batch, ld, Tn);
  }
  return 2;
}

template <class T>
int launch_any(int n, int m, const LqrInT<T> &in, const LqrOutT<T> &out, int *status, T *store,
               int64_t batch, int64_t ld, int Tn, Profiler *prof, cudaStream_t s) {
  if (n != 4) return -1;
  switch (m) {
    case 1: return launch_shape<T, 4, 1>(in, out, status, store, batch, ld, Tn, prof, s);
    case 2: return launch_shape<T, 4, 2>(in, out, status, store, batch, ld, Tn, prof, s);
    case 3: return launch_shape<T, 4, 3>(in, out, status, store, batch, ld, Tn, prof, s);
    case 4: return launch_shape<T, 4, 4>(in, out, status, store, batch, ld, Tn, prof, s);
    default: return -1;
  }
}

}  // namespace

bool f32_supports(int n, int m) { return n == 4 && m >= 1 && m <= 4; }

// The kernels index every array with 32 bits: the largest one (the kept factorization) must
// stay below 2^31 elements.
bool f32_fits_index(int n, int m, int T, int64_t ld) {
  return f32_store_elems(n, m, T) * ld < (int64_t(1) << 31);
}

int64_t f32_store_elems(int n, int m, int T) {
  // the same for every instantiated m: computed from the formula, not from a template
  return int64_t(T + 1) * tri(n) + int64_t(T) * n * m + int64_t(T + 1) * n + int64_t(T) * m;
}

int launch_lqr_factor_solve_f32(int n, int m, const LqrInT<float> &in, const LqrOutT<float> &out,
                                int *status, float *store, int64_t batch, int64_t ld, int T,
                                Profiler *prof, cudaStream_t s) {
  return launch_any<float>(n, m, in, out, status, store, batch, ld, T, prof, s);
}

int launch_lqr_factor_solve_thread_f64(int n, int m, const LqrInT<double> &in,
                                       const LqrOutT<double> &out, int *status, double *store,
                                       int64_t batch, int64_t ld, int T, Profiler *prof,
                                       cudaStream_t s) {
  return launch_any<double>(n, m, in, out, status, store, batch, ld, T, prof, s);
}

}  // namespace sipoc
