// CTA-per-problem Riccati kernels for large state / control dimensions
// (humanoid scale, n = 64, m = 24), sm_100a, FP64.
//
// One CTA of 8 warps owns one problem; every matrix of the stage lives in shared
// memory (column-major, leading dimension == 4 mod 16 so that all FP64 tensor
// core fragment loads are bank-conflict free) and every dense contraction is a
// tiled product on the FP64 tensor cores, `mma.sync.m8n8k4.f64` (DMMA):
// operands are delivered to the MMA as one double per lane, so a 256-FMA
// instruction costs two 8-byte shared-memory loads per lane instead of the
// 2 loads per 4 FMA a register-tiled DFMA product would need (shared memory
// delivers 128 B / clk / SM, DFMA consumes 64 FMA / clk / SM).
//
// Stage algebra (same as riccati_fast.cu; Z = [A_k | B_k | 0], u padded to a
// multiple of 16 with identity in R so the padded G stays positive definite):
//   S    = W' Z                         cta_gemm
//   Psi  = [Q M; M' R] + Z' S           cta_gemm (lower tiles)
//   G^-1 = Psi_uu^-1                    blocked Cholesky + triangular inverse + L^-T L^-1
//   K    = -G^-1 Psi_ux                 cta_gemm                    (lqr.cpp:703-713)
//   V    = Psi_xx + Psi_ux' K           cta_gemm (lower tiles)      (lqr.cpp:715-719)
//   F    = I + D^1/2 V D^1/2 ; W = D^-1/2 (I - F^-1) D^-1/2          (lqr.cpp:475-529)
// Affine part: g = v' - W'(delta' o v' - c'); [w; h] = [q; r] + Z' g;
//   k = -G^-1 h ; v = w + Psi_ux' k.                                (lqr.cpp:778-794)
//
// Global memory is the engine's batch-interleaved layout X[flat * ld + b]; one
// problem's elements are therefore 8-byte accesses ld * 8 bytes apart.  They are
// staged with 8-byte cp.async; neighbouring CTAs (problems b .. b+3) share every
// 32-byte sector through L2, and the path is FP64-bound (25 flop / byte), so the
// sector inefficiency costs L2 bandwidth only.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "riccati_fast.cuh"

namespace sipoc {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kPanelThreads = 32;  // warp 0 runs the Cholesky chain; the others stage

__host__ __device__ constexpr int tri(int n) { return n * (n + 1) / 2; }
__host__ __device__ constexpr int pk(int i, int j, int n) {
  return j * n - j * (j - 1) / 2 + (i - j);
}
__host__ __device__ constexpr int round16(int x) { return (x + 15) / 16 * 16; }
// Smallest leading dimension >= rows with ld % 16 == 4.
__host__ __device__ constexpr int pad_ld(int rows) {
  int ld = rows;
  while (ld % 16 != 4) ++ld;
  return ld;
}

template <int N, int M>
struct CtaSizes {  // per-problem element counts of the store / spill (same as FastSizes)
  static __host__ __device__ constexpr int64_t oW(int) { return 0; }
  static __host__ __device__ constexpr int64_t oK(int T) { return int64_t(T + 1) * tri(N); }
  static __host__ __device__ constexpr int64_t oG(int T) { return oK(T) + int64_t(T) * N * M; }
  static __host__ __device__ constexpr int64_t store(int T) { return oG(T) + int64_t(T) * tri(M); }
  static __host__ __device__ constexpr int64_t ov(int) { return 0; }
  static __host__ __device__ constexpr int64_t ok(int T) { return int64_t(T + 1) * N; }
  static __host__ __device__ constexpr int64_t scratch(int T) { return ok(T) + int64_t(T) * M; }
};

__device__ __forceinline__ void cp_async8(double *smem, const double *gmem) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(double *smem, const double *gmem) {
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
// C(g, 2t), C(g, 2t + 1).
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// Element (i, k) of op(A) for a column-major array with leading dimension ld.
template <bool TRANS>
__device__ __forceinline__ double op_at(const double *A, int ld, int i, int k) {
  return TRANS ? A[i * ld + k] : A[k * ld + i];
}

// (row, column) of the s-th cell of a lower triangle enumerated row by row, s < 36.
__constant__ unsigned char kTriRow[36] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5,
                                          5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 7};
__constant__ unsigned char kTriCol[36] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4, 0, 1, 2,
                                          3, 4, 5, 0, 1, 2, 3, 4, 5, 6, 0, 1, 2, 3, 4, 5, 6, 7};

// One 16 x 16 block (2 x 2 MMA tiles) at (i0, j0) of
//   C (Mo x No) = (ACC ? C : 0) + sign * op(A) (Mo x K) * op(B) (K x No),
// all in shared memory, column-major; executed by one warp.  K a multiple of 4.  Mo and
// No are rounded up to whole 8 x 8 MMA tiles (callers keep that padding zero); tiles
// entirely beyond them are not issued.  KTRI starts the k range at i0 (op(A) upper triangular, i.e. A' of a
// lower-triangular X).
template <bool TA, bool TB, bool ACC, bool KTRI = false>
__device__ __forceinline__ void gemm_block(double *C, int ldc, const double *A, int lda,
                                           const double *B, int ldb, int Mo, int No, int K,
                                           double sign, int i0, int j0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bool r1 = i0 + 8 < Mo, c1 = j0 + 8 < No;
  double acc[2][2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const bool live = (a == 0 || r1) && (b == 0 || c1);
      const double *src = C + (j0 + 8 * b + 2 * t) * ldc + i0 + 8 * a + g;
      acc[a][b][0] = (ACC && live) ? src[0] : 0.0;
      acc[a][b][1] = (ACC && live) ? src[ldc] : 0.0;
    }
  // op(A)(i, k): A column-major -> A[k * lda + i]; transposed -> A[i * lda + k].
  // op(B)(k, j): B column-major -> B[j * ldb + k]; transposed -> B[k * ldb + j].
  const double *pa0 = TA ? A + (i0 + g) * lda + t : A + t * lda + i0 + g;
  const double *pb0 = TB ? B + t * ldb + j0 + g : B + (j0 + g) * ldb + t;
  const int sa = TA ? 1 : lda, sb = TB ? ldb : 1;       // stride of one k step
  const int oa = TA ? 8 * lda : 8, ob = TB ? 8 : 8 * ldb;  // offset of the second MMA tile
  // r1, c1 are warp-uniform: one branch-free k loop per tile shape, and the MMA tiles
  // beyond Mo / No are not issued.
  auto k_loop = [&](auto R1, auto C1) {
#pragma unroll 4
    for (int k0 = KTRI ? i0 : 0; k0 < K; k0 += 4) {
      const double a0 = pa0[k0 * sa];
      const double b0 = sign * pb0[k0 * sb];
      double a1 = 0.0, b1 = 0.0;
      if (decltype(R1)::value) a1 = pa0[k0 * sa + oa];
      if (decltype(C1)::value) b1 = sign * pb0[k0 * sb + ob];
      dmma(acc[0][0], a0, b0);
      if (decltype(C1)::value) dmma(acc[0][1], a0, b1);
      if (decltype(R1)::value) dmma(acc[1][0], a1, b0);
      if (decltype(R1)::value && decltype(C1)::value) dmma(acc[1][1], a1, b1);
    }
  };
  using Yes = std::true_type;
  using No_ = std::false_type;
  if (r1 && c1) k_loop(Yes{}, Yes{});
  else if (r1) k_loop(Yes{}, No_{});
  else if (c1) k_loop(No_{}, Yes{});
  else k_loop(No_{}, No_{});
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      if ((a == 0 || r1) && (b == 0 || c1)) {
        double *dst = C + (j0 + 8 * b + 2 * t) * ldc + i0 + 8 * a + g;
        dst[0] = acc[a][b][0];
        dst[ldc] = acc[a][b][1];
      }
    }
}

// Whole product on the CTA: 16 x 16 blocks dealt round-robin to the warps, slot numbers
// starting at `first` (so that consecutive products of one phase share the deal);
// LOWER computes only blocks on or below the block diagonal (whole blocks).  Returns
// the next free slot number.
template <bool TA, bool TB, bool ACC, bool LOWER, bool KTRI = false>
__device__ int cta_gemm(double *C, int ldc, const double *A, int lda, const double *B, int ldb,
                        int Mo, int No, int K, double sign, int first = 0) {
  const int warp = threadIdx.x >> 5;
  const int tm = (Mo + 15) >> 4, tn = (No + 15) >> 4;
  const int count = LOWER ? tm * (tm + 1) / 2 : tm * tn;
  // first local slot of this warp: smallest s >= 0 with (first + s) % kWarps == warp
  int s = (warp - first) & (kWarps - 1);
  int ti = 0, tj = s;  // (row, column) of slot s for the rectangular deal
  if (!LOWER)
    while (tj >= tn) { tj -= tn; ++ti; }
  for (; s < count; s += kWarps) {
    if (LOWER) { ti = kTriRow[s]; tj = kTriCol[s]; }
    gemm_block<TA, TB, ACC, KTRI>(C, ldc, A, lda, B, ldb, Mo, No, K, sign, ti << 4, tj << 4);
    if (!LOWER) {
      tj += kWarps;
      while (tj >= tn) { tj -= tn; ++ti; }
    }
  }
  return first + count;
}

// y[i] = base[i] + sign * sum_j op(A)(i, j) x[j],  i < rows, j < cols; four threads per
// row (quarters of the columns) reduced with shuffles.  rows <= 64 per pass.
template <bool TRANS>
__device__ void cta_matvec(double *y, const double *base, const double *A, int lda,
                           const double *x, int rows, int cols, double sign) {
  const int tid = threadIdx.x, part = tid & 3;
  for (int r0 = 0; r0 < rows; r0 += kThreads / 4) {
    const int i = r0 + (tid >> 2);
    double acc = 0.0;
    if (i < rows) {
      for (int j = part; j < cols; j += 4) acc += op_at<TRANS>(A, lda, i, j) * x[j];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (i < rows && part == 0) y[i] = (base != nullptr ? base[i] : 0.0) + sign * acc;
  }
}

// Phase timing of the backward kernel (build with -DSIPOC_CTA_TIMING; block 0 prints the
// cycles per phase summed over its stages).  Compiled out otherwise.
#ifdef SIPOC_CTA_TIMING
__device__ unsigned long long g_tick[24];
__shared__ long long tick_last__;
#define TICK(slot)                                         \
  do {                                                     \
    __syncthreads();                                       \
    if (threadIdx.x == 0 && blockIdx.x == 0) {             \
      const long long now__ = clock64();                   \
      atomicAdd(&g_tick[slot], (unsigned long long)(now__ - tick_last__)); \
      tick_last__ = now__;                                 \
    }                                                      \
  } while (0)
// Thread-0-only lap timer (no barrier): for the panel warps' critical path.
#define LAP(slot)                                          \
  do {                                                     \
    if (threadIdx.x == 0 && blockIdx.x == 0) {             \
      const long long now__ = clock64();                   \
      atomicAdd(&g_tick[slot], (unsigned long long)(now__ - lap__)); \
      lap__ = now__;                                       \
    }                                                      \
  } while (0)
#define LAP_BEGIN() long long lap__ = clock64()
#define LAP1(slot)                                         \
  do {                                                     \
    if (threadIdx.x == 32 && blockIdx.x == 0) {            \
      const long long now__ = clock64();                   \
      atomicAdd(&g_tick[slot], (unsigned long long)(now__ - lap__)); \
      lap__ = now__;                                       \
    }                                                      \
  } while (0)
#else
#define TICK(slot) do { } while (0)
#define LAP(slot) do { } while (0)
#define LAP1(slot) do { } while (0)
#define LAP_BEGIN() do { } while (0)
#endif

// Cholesky factor L (packed lower, diagonal included) of the 8 x 8 block at `blk`, in
// registers; d[j] = 1 / L(j, j).  Right-looking, so the dependent chain per column is
// rsqrt -> mul -> one FMA.  Returns false when a pivot is <= 0.
__device__ __forceinline__ bool chol8(const double *blk, int ld, double (&L)[36], double (&d)[8]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i >= j) L[pk(i, j, 8)] = blk[j * ld + i];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double x = L[pk(j, j, 8)];
    ok = ok && (x > 0.0);
    const double r = rsqrt(x);
    d[j] = r;
    L[pk(j, j, 8)] = x * r;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i > j) L[pk(i, j, 8)] *= r;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c > j && i >= c) L[pk(i, c, 8)] -= L[pk(i, j, 8)] * L[pk(c, j, 8)];
    }
  }
  return ok;
}

// In-place inverse of a symmetric positive definite n x n matrix given by its
// LOWER triangle (n a multiple of 8, n <= 64): blocked right-looking Cholesky
// (8-wide panels with look-ahead, DMMA trailing updates), recursive triangular inverse,
// then L^-T L^-1.  On return A holds the inverse: lower triangle always, the full
// symmetric matrix when `mirror`.  `scratch` is an n x ld array, `tbuf` an (n / 2)-column
// array of the same ld, `ddiag` n doubles.  Returns false when a pivot is <= 0 (Eigen
// LLT's failure criterion).  All threads must call it.  `idle(kb, nb)` is called in
// block step kb by the seven warps that are not on the factorization's critical path: the
// caller uses it to issue the next stage's cp.async copies.
// Part 1: the factor.  On return (after a CTA barrier) the lower triangle of A holds L,
// ddiag the reciprocals of its diagonal.
template <class Idle>
__device__ bool cta_cholesky_lookahead(double *A, int ld, int n, double *ddiag, Idle idle,
                                       int n_live = -1) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  bool ok = true;
  const int nb = n >> 3;
  // Rows / columns >= n_live are an identity padding block (their panel rows are exactly
  // zero): the Cholesky leaves them untouched, only their trivial factor is recorded.
  const int nb_live = n_live < 0 ? nb : (n_live + 7) >> 3;
  const int nl = nb_live << 3;
  for (int i = nl + tid; i < n; i += kThreads) ddiag[i] = 1.0;
  TICK(10);
  // Blocked Cholesky with look-ahead.  Warp 0 (the panel warp) owns the critical path:
  // factor the diagonal block kb in registers, solve the panel below it (two rows per
  // lane), update the next diagonal block with that panel, go on to step kb + 1 -- all
  // inside one warp, so the chain only meets __syncwarp.  Warps 1-7 (the trailing
  // warps) apply panel kb first to what the panel warp needs next -- the rest of block
  // column kb + 1 and the diagonal block kb + 2 -- then to the remaining tiles of block
  // columns kb + 2.., then issue staging copies.  Named barriers, alternating with the
  // step parity so that arrivals of consecutive steps never meet on one barrier:
  //   2 / 3  panel kb is complete          (panel warp arrives, trailing warps sync; it
  //          also orders the trailing warps' steps among themselves)
  //   4 / 5  the priority tiles of step kb are done (trailing warps arrive, the panel
  //          warp syncs after the diagonal factor of step kb + 1)
  for (int kb = 0; kb < nb_live; ++kb) {
    const int c0 = kb << 3, c1 = c0 + 8, rem = nl - c1, par = kb & 1;
    if (warp == 0) {
      double L[36], d[8];
      LAP_BEGIN();
      ok = chol8(A + c0 * ld + c0, ld, L, d) && ok;
      LAP(17);
      if (kb >= 1 && rem > 0)  // the trailing warps of step kb - 1 updated this block column
        asm volatile("bar.sync %0, 256;" ::"r"(5 - par) : "memory");
      LAP(18);
      // Slot 0: row `lane` of L21; slot 1: row `lane + 32` of L21 (lanes 0..23), or row
      // `lane - 24` of L11 itself by the same recurrence (lanes 24..31; rem <= 56).
      // x L11' = A(row, c0..c0+8).
      const bool diag_row = lane >= 24;
      int row_idx[2] = {c1 + lane, diag_row ? c0 + (lane - 24) : c1 + 32 + lane};
      bool live[2] = {lane < rem, diag_row || lane + 32 < rem};
      double x[2][8];
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        if (live[z]) {
          const double *row = A + c0 * ld + row_idx[z];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[z][j] = row[j * ld];
        }
      }
      // Right-looking: once x_j is known every later entry takes its term at once, so the
      // dependent chain is one multiply and one FMA per column (same summation order as
      // the row recurrence).
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int z = 0; z < 2; ++z) {
          x[z][j] *= d[j];
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c > j) x[z][c] -= x[z][j] * L[pk(c, j, 8)];
        }
      }
      __syncwarp();  // every lane has read the diagonal block
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        if (live[z]) {
          double *row = A + c0 * ld + row_idx[z];
          const int i = lane - 24;
#pragma unroll
          for (int j = 0; j < 8; ++j) row[j * ld] = (z == 0 || !diag_row || j <= i) ? x[z][j] : 0.0;
        }
      }
      if (diag_row) {  // 1 / L(i, i)
        double di = d[0];
#pragma unroll
        for (int j = 1; j < 8; ++j)
          if (j == lane - 24) di = d[j];
        ddiag[c0 + lane - 24] = di;
      }
      __syncwarp();
      LAP(19);
      if (rem > 8) asm volatile("bar.arrive %0, 256;" ::"r"(2 + par) : "memory");
      if (rem > 0) {
        // A(c1..c1+8, c1..c1+8) -= L21(0..8, :) L21(0..8, :)'
        double *cblk = A + (c1 + 2 * t) * ld + c1 + g;
        double acc[2] = {cblk[0], cblk[ld]};
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const double a = A[(c0 + 4 * s2 + t) * ld + c1 + g];
          dmma(acc, -a, a);
        }
        cblk[0] = acc[0];
        cblk[ld] = acc[1];
        __syncwarp();
        LAP(21);
      }
    } else {
      LAP_BEGIN();
      if (rem > 8) {
        asm volatile("bar.sync %0, 256;" ::"r"(2 + par) : "memory");
        LAP1(12);
        // One 8 x 8 tile C(r0.., q0..) -= L21(r0.., :) L21(q0.., :)' of the trailing matrix.
        auto tile = [&](int r0, int q0) {
          double *cblk = A + (q0 + 2 * t) * ld + r0 + g;
          double acc[2] = {cblk[0], cblk[ld]};
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const double a = -A[(c0 + 4 * s2 + t) * ld + r0 + g];
            const double bb = A[(c0 + 4 * s2 + t) * ld + q0 + g];
            dmma(acc, a, bb);
          }
          cblk[0] = acc[0];
          cblk[ld] = acc[1];
        };
        // Priority tiles: block column kb + 1 below its diagonal block (warps 1..ntile-1)
        // and the diagonal block kb + 2 (warp 7, never needed for the column: ntile <= 7).
        const int ntile = rem >> 3, c2 = c1 + 8, nt2 = ntile - 1;
        if (warp < ntile) tile(c1 + (warp << 3), c1);
        if (warp == kWarps - 1) tile(c2, c2);
        asm volatile("bar.arrive %0, 256;" ::"r"(4 + par) : "memory");
        LAP1(4);
        // The rest of A(c2.., c2..) on 8 x 8 tiles of the lower triangle (slot 0 is the
        // diagonal block done above).
        for (int slot = kWarps - warp; slot < nt2 * (nt2 + 1) / 2; slot += kWarps - 1)
          tile(c2 + (kTriRow[slot] << 3), c2 + (kTriCol[slot] << 3));
        LAP1(22);
      }
      idle(kb, nb_live);
      LAP1(16);
    }
  }
  __syncthreads();
  TICK(11);
  return ok;
}

// Part 2: from the factor in A (lower triangle, reciprocal diagonal in ddiag) to the
// inverse of the matrix, in A.  All threads.
template <int NN>
__device__ void cta_inverse_from_factor(double *A, int ld, double *scratch, double *tbuf,
                                        const double *ddiag, bool mirror) {
  // (the size is a template argument so that the level structure below unrolls into
  // straight-line code: at eight warps per SM every index instruction is exposed latency)
  constexpr int n = NN;
  static_assert(NN % 8 == 0 && NN <= 64, "n is a multiple of 8, at most 64");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  constexpr int nb = n >> 3;
  // X = L^-1 in `scratch`.  Its blocks above the block diagonal are read as zeros.
  for (int blk = warp; blk < 64; blk += kWarps) {  // (bi, bj) over an 8 x 8 grid of blocks
    const int bi = blk & 7, bj = blk >> 3;
    if (bi < bj && bj < nb) {
      double *z = scratch + ((bj << 3) + (lane >> 3)) * ld + (bi << 3) + (lane & 7);
      z[0] = 0.0;
      z[4 * ld] = 0.0;
    }
  }
  // Level 0: the 8 x 8 diagonal blocks.  Lane 8 q + j of a warp computes column j of the
  // inverse of the warp's q-th block by forward substitution on e_j (the entries above
  // the diagonal come out as exact zeros), four blocks per warp.
  for (int kb = warp * 4 + (lane >> 3); kb < nb && warp * 4 < nb; kb += kWarps * 4) {
    const int j = lane & 7;
    const double *blk = A + (kb * 8) * ld + kb * 8;
    double xc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double sacc = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int p = 0; p < 8; ++p)
        if (p < i) sacc -= blk[p * ld + i] * xc[p];
      xc[i] = sacc * ddiag[kb * 8 + i];
    }
    double *dst = scratch + (kb * 8 + j) * ld + kb * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = xc[i];
  }
  __syncthreads();
  TICK(13);
  // Levels s = 8, 16, 32: with L = [L11 0; L21 L22] on 2s x 2s diagonal blocks,
  // X21 = -X22 (L21 X11).  Every level is two batched products on 8 x 8 tiles dealt to
  // all warps (T = L21 X11 into tbuf, then X21), so the dependent chain is
  // 2 log2(n / 8) short products instead of n / 8 block rows.  The blocks of X above
  // the diagonal were zeroed on entry, so every tile runs the full k range (equal trip
  // counts let a warp interleave its tiles).
#pragma unroll
  for (int sz = 8; sz < n; sz <<= 1) {
    const int tc = sz >> 3, lt = sz == 8 ? 0 : (sz == 16 ? 1 : 2);  // tile columns per pair
    const int pairs = (n + 2 * sz - 1) >> (lt + 4);
    const int total = pairs << (2 * lt);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      // Two units per warp at a time, two accumulators per unit (even / odd k steps):
      // four independent DMMA chains of sz / 8 links.
#pragma unroll
      for (int u0 = warp; u0 < total; u0 += 2 * kWarps) {
        const double *pa[2], *pb[2];
        double *dst[2];
        bool live[2];
        int klim[2];  // k range: sz, or the rows the (shorter) last second block really has
#pragma unroll
        for (int z = 0; z < 2; ++z) {
          const int unit = u0 + z * kWarps;
          const int q = unit >> (2 * lt), w = unit & ((1 << (2 * lt)) - 1);
          const int ti = w >> lt, tj = w & (tc - 1);
          const int c0 = 2 * q * sz, r0 = c0 + sz;
          live[z] = unit < total && r0 + (ti << 3) < n;
          klim[z] = (pass == 1 && n - r0 < sz) ? n - r0 : sz;
          double *tq = tbuf + (q * sz) * ld;  // T of pair q: (rows of L21) x sz
          if (pass == 0) {  // T = L21 X11
            pa[z] = A + (c0 + t) * ld + r0 + (ti << 3) + g;
            pb[z] = scratch + (c0 + (tj << 3) + g) * ld + c0 + t;
            dst[z] = tq + ((tj << 3) + 2 * t) * ld + (ti << 3) + g;
          } else {  // X21 = -X22 T
            pa[z] = scratch + (r0 + t) * ld + r0 + (ti << 3) + g;
            pb[z] = tq + ((tj << 3) + g) * ld + t;
            dst[z] = scratch + (c0 + (tj << 3) + 2 * t) * ld + r0 + (ti << 3) + g;
          }
        }
        double acc[2][2][2] = {};
#pragma unroll
        for (int k0 = 0; k0 < sz; k0 += 8) {
#pragma unroll
          for (int z = 0; z < 2; ++z) {
            if (live[z] && k0 < klim[z]) {  // (klim is a multiple of 8)
              const double a0 = pa[z][k0 * ld], a1 = pa[z][(k0 + 4) * ld];
              const double b0 = pb[z][k0], b1 = pb[z][k0 + 4];
              dmma(acc[z][0], a0, b0);
              dmma(acc[z][1], a1, b1);
            }
          }
        }
        const double sgn = pass == 0 ? 1.0 : -1.0;
#pragma unroll
        for (int z = 0; z < 2; ++z) {
          if (live[z]) {
            dst[z][0] = sgn * (acc[z][0][0] + acc[z][1][0]);
            dst[z][ld] = sgn * (acc[z][0][1] + acc[z][1][1]);
          }
        }
      }
      __syncthreads();
    }
  }
  TICK(14);
  // A = L^-T L^-1 (lower blocks, whole diagonal blocks): (i, j) sums over k >= i only.
  cta_gemm<true, false, false, true, true>(A, ld, scratch, ld, scratch, ld, n, n, n, 1.0);
  __syncthreads();
  TICK(15);
  if (mirror) {
    for (int e = tid; e < n * n; e += kThreads) {
      const int i = e % n, j = e / n;
      if (i > j) A[i * ld + j] = A[j * ld + i];
    }
    __syncthreads();
  }
  TICK(16);
}

template <int NN, class Idle>
__device__ bool cta_spd_inverse(double *A, int ld, double *scratch, double *tbuf, double *ddiag,
                                Idle idle, int n_live = -1, bool mirror = true) {
  const bool ok = cta_cholesky_lookahead(A, ld, NN, ddiag, idle, n_live);
  cta_inverse_from_factor<NN>(A, ld, scratch, tbuf, ddiag, mirror);
  return ok;
}

// Cholesky of a small matrix (n_live <= 32) by ONE warp, no barrier but __syncwarp: the
// same 8-wide panels, every tile of the trailing update done by the warp itself.  Used
// for G, whose factorization then runs beside the Psi products of the other warps.
__device__ bool warp_cholesky(double *A, int ld, int n, double *ddiag, int n_live) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nl = ((n_live + 7) >> 3) << 3;
  bool ok = true;
  for (int i = nl + lane; i < n; i += 32) ddiag[i] = 1.0;
  for (int c0 = 0; c0 < nl; c0 += 8) {
    const int c1 = c0 + 8, rem = nl - c1;  // rem <= 24
    double L[36], d[8];
    ok = chol8(A + c0 * ld + c0, ld, L, d) && ok;
    // lanes < rem: a row of L21; lanes 24..31: the rows of L11 itself (same recurrence)
    const bool diag_row = lane >= 24;
    const bool live = diag_row || lane < rem;
    double *row = A + c0 * ld + (diag_row ? c0 + lane - 24 : c1 + lane);
    double x[8];
    if (live) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = row[j * ld];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[j] *= d[j];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c > j) x[c] -= x[j] * L[pk(c, j, 8)];
    }
    __syncwarp();  // every lane has read the diagonal block
    if (live) {
#pragma unroll
      for (int j = 0; j < 8; ++j) row[j * ld] = (!diag_row || j <= lane - 24) ? x[j] : 0.0;
    }
    if (lane == 23) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ddiag[c0 + j] = d[j];
    }
    __syncwarp();
    // trailing update A(c1.., c1..) -= L21 L21' on the lower 8 x 8 tiles
    for (int ti = 0; ti < (rem >> 3); ++ti)
      for (int tj = 0; tj <= ti; ++tj) {
        const int r0 = c1 + (ti << 3), q0 = c1 + (tj << 3);
        double *cblk = A + (q0 + 2 * t) * ld + r0 + g;
        double acc[2] = {cblk[0], cblk[ld]};
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const double a = -A[(c0 + 4 * s2 + t) * ld + r0 + g];
          const double bb = A[(c0 + 4 * s2 + t) * ld + q0 + g];
          dmma(acc, a, bb);
        }
        cblk[0] = acc[0];
        cblk[ld] = acc[1];
      }
    __syncwarp();
  }
  return ok;
}

// Even deal of the lower triangle of an N x N matrix over the CTA: columns j and N - 1 - j
// together hold N + 1 entries, so the N / 2 column pairs are equal work; kThreads / (N / 2)
// threads share a pair.  `load(i, j)` is called for all of a thread's entries first, then
// `store(i, j, value)` -- the loads of one pass are in flight together.
template <int N, class Load, class Store>
__device__ __forceinline__ void for_lower_triangle(Load load, Store store) {
  constexpr int kPairs = N / 2, kPer = kThreads / kPairs, kMax = (N + 1 + kPer - 1) / kPer;
  static_assert(N % 2 == 0 && kThreads % kPairs == 0, "column pairs divide the CTA");
  const int p = threadIdx.x / kPer, sub = threadIdx.x % kPer;
  double val[kMax];
#pragma unroll
  for (int u = 0; u < kMax; ++u) {
    const int idx = sub + u * kPer;
    if (idx <= N) {
      const bool first = idx < N - p;
      const int j = first ? p : N - 1 - p;
      const int i = first ? p + idx : idx - 1;  // second column: N - 1 - p + (idx - (N - p))
      val[u] = load(i, j);
    }
  }
#pragma unroll
  for (int u = 0; u < kMax; ++u) {
    const int idx = sub + u * kPer;
    if (idx <= N) {
      const bool first = idx < N - p;
      const int j = first ? p : N - 1 - p;
      const int i = first ? p + idx : idx - 1;
      store(i, j, val[u]);
    }
  }
}

// Shared-memory map (doubles).
template <int N, int M>
struct CtaSmem {
  static constexpr int MP = round16(M);
  static constexpr int NZ = N + MP;
  static constexpr int LDN = pad_ld(N);    // matrices with N rows
  static constexpr int LDM = pad_ld(MP);   // matrices with MP rows
  static constexpr int oW = 0;                        // W' -> Psi_xx -> F -> W   (N x N)
  static constexpr int oZ = oW + N * LDN;             // Z = [A | B | 0]          (N x NZ)
  static constexpr int oS = oZ + NZ * LDN;            // S = W' Z; later scratch  (N x NZ)
  static constexpr int oPux = oS + NZ * LDN;          // Psi_ux                   (MP x N)
  static constexpr int oPuu = oPux + N * LDM;         // Psi_uu -> G^-1           (MP x MP)
  static constexpr int oK = oPuu + MP * LDM;          // K                        (MP x N)
  static constexpr int oGs = oK + N * LDM;            // scratch of the G inverse (MP x MP)
  static constexpr int oDd = oGs + MP * LDM;          // 1 / L(j, j) of the factor in flight
  static_assert(2 * MP >= N, "the F inverse parks its T blocks in columns N.. of S");
  static constexpr int oVec = oDd + N;
  static constexpr int vq = oVec, vr = vq + N, vc = vr + MP, vd = vc + N, vv = vd + N,
                       vdl = vv + N, vsd = vdl + N, vsdi = vsd + N, vf = vsdi + N, vg = vf + N,
                       vhw = vg + N, vkk = vhw + NZ, vEnd = vkk + MP;
  static constexpr int kDoubles = vEnd;
  static constexpr int kBytes = kDoubles * int(sizeof(double));
};

template <int N, int M, bool SOLVE>
__global__ void __launch_bounds__(kThreads, 1)
riccati_backward_cta(LqrIn pm, int *status_out, double *store, double *scratch, int64_t batch,
                     int64_t ld, int T) {
  using S = CtaSmem<N, M>;
  using Zs = CtaSizes<N, M>;
  constexpr int MP = S::MP, LDN = S::LDN, LDM = S::LDM;
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  double *Wp = sm + S::oW, *Zb = sm + S::oZ, *Sb = sm + S::oS, *Pux = sm + S::oPux,
         *Puu = sm + S::oPuu, *Kb = sm + S::oK, *Gs = sm + S::oGs,
         *Dd = sm + S::oDd;
  double *q_s = sm + S::vq, *r_s = sm + S::vr, *c_s = sm + S::vc, *d_s = sm + S::vd,
         *v_s = sm + S::vv, *dl_s = sm + S::vdl, *sd_s = sm + S::vsd, *sdi_s = sm + S::vsdi,
         *f_s = sm + S::vf, *g_s = sm + S::vg, *hw_s = sm + S::vhw, *kk_s = sm + S::vkk;

  // The store (W, K, G^-1) and the spill (v, k) are private to this file's kernels and
  // problem-major like the inputs: every store below is a contiguous run.
  double *Wst = store + static_cast<size_t>(b) * Zs::store(T) + Zs::oW(T);
  double *Kst = store + static_cast<size_t>(b) * Zs::store(T) + Zs::oK(T);
  double *Gst = store + static_cast<size_t>(b) * Zs::store(T) + Zs::oG(T);
  double *vst = SOLVE ? scratch + static_cast<size_t>(b) * Zs::scratch(T) + Zs::ov(T) : nullptr;
  double *kst = SOLVE ? scratch + static_cast<size_t>(b) * Zs::scratch(T) + Zs::ok(T) : nullptr;

  // Inputs come from the problem-major copies [problem][flat]: this CTA's problem is a
  // contiguous run per array, so a warp's 32 eight-byte copies cover 8 whole sectors
  // (the batch-interleaved layout would make them 32 sectors with 8 useful bytes each).
  const size_t ub = static_cast<size_t>(b);
  const double *bQ = pm.Q + ub * (static_cast<size_t>(T + 1) * N * N);
  const double *bM = pm.M + ub * (static_cast<size_t>(T) * N * M);
  const double *bR = pm.R + ub * (static_cast<size_t>(T) * M * M);
  const double *bA = pm.A + ub * (static_cast<size_t>(T) * N * N);
  const double *bB = pm.B + ub * (static_cast<size_t>(T) * N * M);
  const double *bd = pm.delta + ub * (static_cast<size_t>(T + 1) * N);
  const double *bq = SOLVE ? pm.q + ub * (static_cast<size_t>(T + 1) * N) : nullptr;
  const double *bc = SOLVE ? pm.c + ub * (static_cast<size_t>(T + 1) * N) : nullptr;
  const double *br = SOLVE ? pm.r + ub * (static_cast<size_t>(T) * M) : nullptr;

  // Staging copies.  (t0, nt): the calling threads are tid in [t0, t0 + nt); (part, parts):
  // this call moves the part-th of `parts` equal slices of the element range.
  auto stage_edge_z = [&](int k, int t0, int nt, int part, int parts) {  // A_k, B_k -> Z
    const double *gA = bA + static_cast<size_t>(k) * N * N;
    const double *gB = bB + static_cast<size_t>(k) * N * M;
    // Columns of A and B are contiguous both in the problem-major copy and in Z: pairs
    // of rows move as 16-byte copies.
    static_assert(N % 2 == 0 && LDN % 2 == 0 && S::oZ % 2 == 0, "16-byte staging of Z");
    constexpr int total = (N * N + N * M) / 2;
    const int per = (total + parts - 1) / parts;
    const int hi = (part + 1) * per < total ? (part + 1) * per : total;
    for (int p = part * per + (tid - t0); p < hi; p += nt) {
      const int e = 2 * p;
      if (e < N * N) {
        cp_async16(Zb + (e / N) * LDN + e % N, gA + e);
      } else {
        const int f = e - N * N;
        cp_async16(Zb + (N + f / N) * LDN + f % N, gB + f);
      }
    }
  };
  // M_k' -> Psi_ux, R_k (lower) -> Psi_uu, vectors of stage k
  auto stage_edge_rest = [&](int k, int t0, int nt) {
    const int t = tid - t0;
    const double *gM = bM + static_cast<size_t>(k) * N * M;
    for (int e = t; e < N * M; e += nt)
      cp_async8(Pux + (e % N) * LDM + e / N, gM + e);  // M(x, u) -> Psi_ux(u, x)
    const double *gR = bR + static_cast<size_t>(k) * M * M;
    for (int e = t; e < M * M; e += nt)
      if (e % M >= e / M) cp_async8(Puu + (e / M) * LDM + e % M, gR + e);
    for (int i = t; i < N; i += nt) {
      cp_async8(d_s + i, bd + static_cast<size_t>(k) * N + i);
      if (SOLVE) {
        cp_async8(q_s + i, bq + static_cast<size_t>(k) * N + i);
        cp_async8(c_s + i, bc + static_cast<size_t>(k + 1) * N + i);
      }
    }
    if (SOLVE)
      for (int a = t; a < M; a += nt)
        cp_async8(r_s + a, br + static_cast<size_t>(k) * M + a);
  };
  // Q_k (lower, packed) -> the K buffer, which is idle between the end of one stage and
  // the K product of the next; unpacked into Psi_xx once W' has been consumed.
  static_assert(N * LDM >= tri(N), "packed Q must fit the K buffer");
  auto stage_q_packed = [&](int k, int t0, int nt) {
    const double *gQ = bQ + static_cast<size_t>(k) * N * N;
    for (int e = tid - t0; e < N * N; e += nt) {
      const int i = e % N, j = e / N;
      if (i >= j) cp_async8(Kb + pk(i, j, N), gQ + e);
    }
  };
  auto stage_q_lower = [&](int k) {  // Q_k (lower) -> Psi_xx buffer
    const double *gQ = bQ + static_cast<size_t>(k) * N * N;
    for (int e = tid; e < N * N; e += kThreads)
      if (e % N >= e / N) cp_async8(Wp + (e / N) * LDN + e % N, gQ + e);
  };

  int status = SIPOC_FACTOR_SUCCESS;

  // Node processing on the N x N matrix in Wp (lower triangle = V_k), v_k in hw_s[0..N):
  // delta check, F, inverse, W (full, in Wp), stores W_k and v_k.
  auto process_node = [&](int k) {
    bool d_ok = true;
    for (int i = tid; i < N; i += kThreads) {
      const double d = d_s[i];
      d_ok = d_ok && (d > 0.0);
      const double sd = sqrt(d);
      dl_s[i] = d;
      sd_s[i] = sd;
      sdi_s[i] = 1.0 / sd;
      if (SOLVE) {
        const double v = hw_s[i];
        v_s[i] = v;
        vst[static_cast<size_t>(k) * N + i] = v;
      }
    }
    const bool any_bad = __syncthreads_or(!d_ok);
    if (any_bad && status == SIPOC_FACTOR_SUCCESS) status = SIPOC_FACTOR_INVALID_DELTA;
    // F = I + D^1/2 V D^1/2 on the lower triangle.
    for_lower_triangle<N>([&](int i, int j) { return Wp[j * LDN + i]; },
                          [&](int i, int j, double v) {
                            Wp[j * LDN + i] = sd_s[i] * v * sd_s[j] + (i == j ? 1.0 : 0.0);
                          });
    __syncthreads();
    TICK(8);
    // M, R, q, r, c, delta of the next stage (and, after the terminal node, its A and B)
    // are issued by the idle warps of the F factorization's block steps.
    const bool f_ok = cta_spd_inverse<N>(Wp, LDN, Sb, Sb + N * LDN, Dd, [&](int kb, int nb) {
      if (k == 0) return;
      if (kb == 0) stage_edge_rest(k - 1, kPanelThreads, kThreads - kPanelThreads);
      if (kb == 1) stage_q_packed(k - 1, kPanelThreads, kThreads - kPanelThreads);
      if (k == T && kb >= 1 && kb < nb) stage_edge_z(k - 1, kPanelThreads, kThreads - kPanelThreads, kb - 1, nb - 1);
      cp_async_commit();
    }, -1, false);  // the W pass below reads the lower triangle only
    if (!__syncthreads_and(f_ok) && status == SIPOC_FACTOR_SUCCESS)
      status = SIPOC_FACTOR_F_FACTORIZATION_FAILURE;
    // W = D^-1/2 (I - F^-1) D^-1/2: both triangles in shared memory, packed lower to the store.
    for_lower_triangle<N>([&](int i, int j) { return Wp[j * LDN + i]; },
                          [&](int i, int j, double v) {
                            const double w = sdi_s[i] * ((i == j ? 1.0 : 0.0) - v) * sdi_s[j];
                            Wp[j * LDN + i] = w;
                            Wp[i * LDN + j] = w;
                            // the store keeps P = F^-1: the rollout applies (I + D V)^-1 as
                            // D^1/2 P D^-1/2 (lqr.cpp:531-549), which does not cancel at large delta
                            Wst[static_cast<size_t>(k) * tri(N) + pk(i, j, N)] = v;
                          });
    __syncthreads();
    TICK(9);
  };

  // Padding, written once: columns N+M.. of Z are zero, rows M.. of Psi_ux are zero,
  // Psi_uu carries identity on its padded diagonal (blkdiag(G, I) keeps that shape
  // through the inverse), r is zero beyond M.
  for (int e = tid; e < (MP - M) * N; e += kThreads) Zb[(N + M + e / N) * LDN + e % N] = 0.0;
  for (int e = tid; e < N * LDM; e += kThreads) Pux[e] = 0.0;
  for (int e = tid; e < MP * LDM; e += kThreads) {
    const int i = e % LDM, j = e / LDM;
    Puu[e] = (i == j && i >= M && i < MP) ? 1.0 : 0.0;
  }
  if (tid < MP) r_s[tid] = 0.0;
#ifdef SIPOC_CTA_TIMING
  if (tid == 0 && b == 0) {
    for (int i = 0; i < 23; ++i) g_tick[i] = 0;
    tick_last__ = clock64();
  }
#endif
  __syncthreads();

  // ---- terminal node ----------------------------------------------------------
  stage_q_lower(T);
  if (tid < N) {
    cp_async8(d_s + tid, bd + static_cast<size_t>(T) * N + tid);
    if (SOLVE) cp_async8(hw_s + tid, bq + static_cast<size_t>(T) * N + tid);
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  process_node(T);

  for (int k = T - 1; k >= 0; --k) {
    TICK(0);
    cp_async_wait_all();
    __syncthreads();
    TICK(1);

    if (SOLVE) {
      // g = v' - W'(delta' o v' - c')
      for (int i = tid; i < N; i += kThreads) f_s[i] = dl_s[i] * v_s[i] - c_s[i];
      __syncthreads();
      cta_matvec<false>(g_s, v_s, Wp, LDN, f_s, N, N, -1.0);
      __syncthreads();
      // [w; h] = [q; r] + Z' g
      cta_matvec<true>(hw_s, q_s, Zb, LDN, g_s, N, N, 1.0);
      cta_matvec<true>(hw_s + N, r_s, Zb + N * LDN, LDN, g_s, MP, N, 1.0);
    }
    TICK(2);
    // S = W' Z
    cta_gemm<false, false, false, false>(Sb, LDN, Wp, LDN, Zb, LDN, N, N + M, N, 1.0);
    __syncthreads();
    TICK(3);
    // Psi_xx base: Q_k lower (prefetched, packed, in the K buffer) into the now free W' buffer.
    for_lower_triangle<N>([&](int i, int j) { return Kb[pk(i, j, N)]; },
                          [&](int i, int j, double v) { Wp[j * LDN + i] = v; });
    __syncthreads();
    // Psi_uu += B' S_u first (three blocks, warps 0-2); as soon as it is complete warp 0
    // factors G = Psi_uu on its own (warp_cholesky) while warps 1-7 form Psi_ux += B' S_x
    // and Psi_xx += A' S_x (lower blocks): the 24 pivots of G run beside the two big
    // products instead of after them.
    constexpr int kTmU = (M + 15) / 16, kTnX = N / 16;
    constexpr int kPuuBlocks = kTmU * (kTmU + 1) / 2, kPuxBlocks = kTmU * kTnX;
    constexpr int kPsiBlocks = kPuxBlocks + kTnX * (kTnX + 1) / 2;
    static_assert(kPuuBlocks <= 3, "Psi_uu is dealt to warps 0-2");
    const int warp = tid >> 5;
    bool g_chol_ok = true;
    if (warp < 3) {
      if (warp < kPuuBlocks)
        gemm_block<true, false, true>(Puu, LDM, Zb + N * LDN, LDN, Sb + N * LDN, LDN, M, M, N, 1.0,
                                      kTriRow[warp] << 4, kTriCol[warp] << 4);
      asm volatile("bar.sync 8, 96;" ::: "memory");
    }
    if (warp == 0) {
      g_chol_ok = warp_cholesky(Puu, LDM, MP, Dd, M);
    } else {
      // Blocks in service order: warps 3-7 take the first five (warps 1-2 are still on
      // Psi_uu), then 3-7, 1, 2 in turn.
      for (int blk = 0; blk < kPsiBlocks; ++blk) {
        const int r = blk < 5 ? blk : (blk - 5) % 7;
        const int owner = r < 5 ? 3 + r : r - 4;
        if (owner != warp) continue;
        if (blk < kPuxBlocks) {
          gemm_block<true, false, true>(Pux, LDM, Zb + N * LDN, LDN, Sb, LDN, M, N, N, 1.0,
                                        (blk / kTnX) << 4, (blk % kTnX) << 4);
        } else {
          const int s2 = blk - kPuxBlocks;
          gemm_block<true, false, true>(Wp, LDN, Zb, LDN, Sb, LDN, N, N, N, 1.0, kTriRow[s2] << 4,
                                        kTriCol[s2] << 4);
        }
      }
      // Z is consumed once all seven warps are through: A, B of the next stage fly during
      // the rest of this one.
      asm volatile("bar.sync 9, %0;" ::"r"(kThreads - 32) : "memory");
      if (k > 0) {
        stage_edge_z(k - 1, 32, kThreads - 32, 0, 1);
        cp_async_commit();
      }
    }
    const bool g_ok = __syncthreads_and(g_chol_ok);
    TICK(5);
    // G^-1 (full, in Puu) from its factor.
    cta_inverse_from_factor<(M + 7) / 8 * 8>(Puu, LDM, Gs, Sb, Dd, true);  // the live part: the padding block stays I
    if (!g_ok && status == SIPOC_FACTOR_SUCCESS) status = SIPOC_FACTOR_G_FACTORIZATION_FAILURE;
    TICK(6);
    // K = -G^-1 Psi_ux
    // (rows and the k range stop at M: the padding block of G^-1 is the identity and the
    // padding rows of Psi_ux are zero)
    cta_gemm<false, false, false, false>(Kb, LDM, Puu, LDM, Pux, LDM, M, N, M, -1.0);
    if (SOLVE) cta_matvec<false>(kk_s, nullptr, Puu, LDM, hw_s + N, MP, MP, -1.0);  // k = -G^-1 h
    __syncthreads();
    // V = Psi_xx + Psi_ux' K (lower blocks);  v = w + Psi_ux' k
    cta_gemm<true, false, true, true>(Wp, LDN, Pux, LDM, Kb, LDM, N, N, M, 1.0);
    if (SOLVE) cta_matvec<true>(hw_s, hw_s, Pux, LDM, kk_s, N, MP, 1.0);
    // stores of the edge: K (M x N), G^-1 (packed lower), k
    for (int e = tid; e < N * M; e += kThreads)
      Kst[static_cast<size_t>(k) * N * M + e] = Kb[(e / M) * LDM + e % M];
    for (int e = tid; e < M * M; e += kThreads) {
      const int i = e % M, j = e / M;
      if (i >= j)
        Gst[static_cast<size_t>(k) * tri(M) + pk(i, j, M)] = Puu[j * LDM + i];
    }
    if (SOLVE && tid < M) kst[static_cast<size_t>(k) * M + tid] = kk_s[tid];
    __syncthreads();
    TICK(7);
    process_node(k);
  }
#ifdef SIPOC_CTA_TIMING
  __syncthreads();
  if (tid == 0 && b == 0) {
    __threadfence();
    // 0 loop top, 1 staging wait, 2 affine, 3 S, 4 u-block, 5 Psi_xx, 6 G inverse (10..16 inside
    // both inverses), 7 K / V / stores, 8 F build, 9 W finish; inside the inverses: 10 entry,
    // 11 diagonal block + panel, 12 trailing update, 13 block inverses, 14 L^-1, 15 L^-T L^-1,
    // 16 mirror.
    for (int i = 0; i < 23; ++i) printf("tick %2d  %10llu cycles / stage\n", i, *(volatile unsigned long long *)&g_tick[i] / T);
  }
#endif
  if (status_out != nullptr && tid == 0) status_out[b] = status;
}

// Backward affine sweep against a kept factorization, one CTA per problem; inputs from
// the problem-major copies, W / K / G^-1 from the problem-major store.
template <int N, int M>
__global__ void __launch_bounds__(kThreads)
affine_backward_cta(LqrIn pm, const double *store, double *scratch, int64_t batch, int64_t ld,
                    int T) {
  using Zs = CtaSizes<N, M>;
  __shared__ double v[N], f[N], g[N], h[M], kk[M], sdi[N];
  const int tid = threadIdx.x;
  const size_t b = blockIdx.x;
  const double *Wst = store + b * Zs::store(T) + Zs::oW(T);
  const double *Kst = store + b * Zs::store(T) + Zs::oK(T);
  const double *Gst = store + b * Zs::store(T) + Zs::oG(T);
  double *vst = scratch + b * Zs::scratch(T) + Zs::ov(T);
  double *kst = scratch + b * Zs::scratch(T) + Zs::ok(T);
  const double *gq = pm.q + b * (static_cast<size_t>(T + 1) * N);
  const double *gc = pm.c + b * (static_cast<size_t>(T + 1) * N);
  const double *gd = pm.delta + b * (static_cast<size_t>(T + 1) * N);
  const double *gr = pm.r + b * (static_cast<size_t>(T) * M);
  const double *gA = pm.A + b * (static_cast<size_t>(T) * N * N);
  const double *gB = pm.B + b * (static_cast<size_t>(T) * N * M);
  for (int i = tid; i < N; i += kThreads) {
    v[i] = gq[T * N + i];
    vst[T * N + i] = v[i];
  }
  __syncthreads();
  const int row = tid >> 2, part = tid & 3;  // 4 threads per output row (N <= 64)
  for (int k = T - 1; k >= 0; --k) {
    for (int i = tid; i < N; i += kThreads) {
      const double d = gd[(k + 1) * N + i];
      sdi[i] = rsqrt(d);
      f[i] = sdi[i] * (d * v[i] - gc[(k + 1) * N + i]);  // s = D^-1/2 (delta' o v' - c')
    }
    __syncthreads();
    {  // g = v - W' (delta' o v' - c'),  W' z = D^-1/2 (s - P s) with the kept P = F^-1
      double acc = 0.0;
      if (row < N)
        for (int j = part; j < N; j += 4) {
          const int i = row;
          acc += __ldg(Wst + (k + 1) * tri(N) + (i >= j ? pk(i, j, N) : pk(j, i, N))) * f[j];
        }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (row < N && part == 0) g[row] = v[row] - sdi[row] * (f[row] - acc);
    }
    __syncthreads();
    {  // h = r + B' g
      double acc = 0.0;
      if (row < M)
        for (int p = part; p < N; p += 4) acc += __ldg(gB + (k * M + row) * N + p) * g[p];
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (row < M && part == 0) h[row] = gr[k * M + row] + acc;
    }
    __syncthreads();
    {  // k = -G^-1 h
      double acc = 0.0;
      if (row < M)
        for (int j = part; j < M; j += 4) {
          const int i = row;
          acc += __ldg(Gst + k * tri(M) + (i >= j ? pk(i, j, M) : pk(j, i, M))) * h[j];
        }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (row < M && part == 0) {
        kk[row] = -acc;
        kst[k * M + row] = -acc;
      }
    }
    __syncthreads();
    {  // v = q + A' g + K' h
      double acc = 0.0;
      if (row < N) {
        for (int p = part; p < N; p += 4) acc += __ldg(gA + (k * N + row) * N + p) * g[p];
        for (int a = part; a < M; a += 4) acc += __ldg(Kst + (k * N + row) * M + a) * h[a];
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      __syncthreads();
      if (row < N && part == 0) {
        v[row] = gq[k * N + row] + acc;
        vst[k * N + row] = v[row];
      }
    }
    __syncthreads();
  }
}

// ---- mbarrier + bulk async copy (TMA, 1-D) --------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (done == 0);
}
// bytes: a multiple of 16; both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes,
                                         unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Per-stage operand block of the rollout in shared memory (doubles).
template <int N, int M>
struct RolloutStage {
  static constexpr int oW = 0;                 // W_{k+1}, packed lower
  static constexpr int oK = oW + tri(N);       // K_k   (M x N)
  static constexpr int oA = oK + N * M;        // A_k   (N x N)
  static constexpr int oB = oA + N * N;        // B_k   (N x M)
  static constexpr int ov = oB + N * M;        // v_{k+1}
  static constexpr int od = ov + N;            // delta_{k+1}
  static constexpr int oc = od + N;            // c_{k+1}
  static constexpr int ok = oc + N;            // k_k
  static constexpr int kDoubles = ok + round16(M);
  static_assert(tri(N) % 2 == 0 && (N * M) % 2 == 0 && N % 2 == 0 && M % 2 == 0,
                "every block is a whole number of 16-byte units");
};

// Root solve + forward rollout + costates (lqr.cpp:798-870), one CTA per problem.  The
// rollout is a chain of small matrix-vector products whose operands are used once: it
// runs at the speed its operands arrive.  Everything a stage needs is contiguous in the
// problem-major store / spill / input copies, so one thread fetches a stage with eight
// bulk async copies (TMA) into one of two shared-memory blocks and an mbarrier reports
// their arrival; the fetch of stage k + 1 flies while stage k is computed.
template <int N, int M>
__global__ void __launch_bounds__(kThreads)
rollout_forward_cta(LqrIn pm, LqrOut out, const double *store, const double *scratch,
                    int64_t batch, int64_t ld, int T) {
  static_assert(kThreads % N == 0 && M <= N && M <= 32, "kThreads / N threads per row");
  using Zs = CtaSizes<N, M>;
  using St = RolloutStage<N, M>;
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ double x[N], u[round16(M)], f[N], ax[N], sdi[N], red[kThreads], red2[kThreads];
  const int tid = threadIdx.x;
  const size_t b = blockIdx.x;
  const size_t L_ = static_cast<size_t>(ld);
  const double *Wst = store + b * Zs::store(T) + Zs::oW(T);
  const double *Kst = store + b * Zs::store(T) + Zs::oK(T);
  const double *vst = scratch + b * Zs::scratch(T) + Zs::ov(T);
  const double *kst = scratch + b * Zs::scratch(T) + Zs::ok(T);
  const double *gA = pm.A + b * (static_cast<size_t>(T) * N * N);
  const double *gB = pm.B + b * (static_cast<size_t>(T) * N * M);
  const double *gc = pm.c + b * (static_cast<size_t>(T + 1) * N);
  const double *gd = pm.delta + b * (static_cast<size_t>(T + 1) * N);
  double *xo = out.x + b, *uo = out.u + b, *yo = out.y + b;

  // fetch(k, slot): operands of edge k (node k + 1); k = -1 fetches the root node only.
  auto fetch = [&](int k, int slot) {
    double *dst = sm + slot * St::kDoubles;
    const unsigned vec = N * sizeof(double);
    unsigned bytes = tri(N) * sizeof(double) + 3 * vec;
    if (k >= 0) bytes += (2 * N * M + N * N + M) * sizeof(double);
    mbar_expect_tx(&bar[slot], bytes);
    bulk_g2s(dst + St::oW, Wst + static_cast<size_t>(k + 1) * tri(N), tri(N) * sizeof(double),
             &bar[slot]);
    bulk_g2s(dst + St::ov, vst + static_cast<size_t>(k + 1) * N, vec, &bar[slot]);
    bulk_g2s(dst + St::od, gd + static_cast<size_t>(k + 1) * N, vec, &bar[slot]);
    bulk_g2s(dst + St::oc, gc + static_cast<size_t>(k + 1) * N, vec, &bar[slot]);
    if (k >= 0) {
      bulk_g2s(dst + St::oK, Kst + static_cast<size_t>(k) * N * M, N * M * sizeof(double),
               &bar[slot]);
      bulk_g2s(dst + St::oA, gA + static_cast<size_t>(k) * N * N, N * N * sizeof(double),
               &bar[slot]);
      bulk_g2s(dst + St::oB, gB + static_cast<size_t>(k) * N * M, N * M * sizeof(double),
               &bar[slot]);
      bulk_g2s(dst + St::ok, kst + static_cast<size_t>(k) * M, M * sizeof(double), &bar[slot]);
    }
  };

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    fetch(-1, 0);
    if (T > 0) fetch(0, 1);
  }
  // Thread (part, row): row = tid % N, part = tid / N.  A warp reads one column of an
  // operand at consecutive rows (conflict-free); the P partial sums of a row meet in `red`.
  constexpr int P = kThreads / N;
  const int row = tid % N, part = tid / N;
  // (W f)(row), this part's columns, from the packed lower triangle
  auto w_times_f = [&](const double *Wk) {
    double acc = 0.0;
#pragma unroll 4
    for (int j = part; j < N; j += P) acc += Wk[row >= j ? pk(row, j, N) : pk(j, row, N)] * f[j];
    return acc;
  };
  auto total = [&](const double *r, int i) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < P; ++q) acc += r[q * N + i];
    return acc;
  };

  // The store keeps P = F^-1 per node.  With s = D^-1/2 f and t = P s:
  //   (I + D V)^-1 f = D^1/2 t (lqr.cpp:531-549),  W f = D^-1/2 (s - t).
  // ---- root: f = delta v - c,  x = -(I + D V)^-1 f,  y = v - W f
  mbar_wait(&bar[0], 0);
  {
    const double *blk = sm;
    if (tid < N) {
      const double d = blk[St::od + tid];
      sdi[tid] = rsqrt(d);
      f[tid] = sdi[tid] * (d * blk[St::ov + tid] - blk[St::oc + tid]);
    }
    __syncthreads();
    red[part * N + row] = w_times_f(blk + St::oW);
    __syncthreads();
    if (tid < N) {
      const double t = total(red, tid);
      const double xi = -blk[St::od + tid] * sdi[tid] * t;
      x[tid] = xi;
      __stcs(xo + static_cast<size_t>(tid) * L_, xi);
      __stcs(yo + static_cast<size_t>(tid) * L_, blk[St::ov + tid] - sdi[tid] * (f[tid] - t));
    }
    __syncthreads();  // slot 0 is free, x is visible
  }
  for (int k = 0; k < T; ++k) {
    const int slot = (k + 1) & 1;  // stage k sits in slot (k + 1) % 2; the root used slot 0
    if (tid == 0 && k + 1 < T) fetch(k + 1, slot ^ 1);
    mbar_wait(&bar[slot], ((k + 1) >> 1) & 1);
    const double *blk = sm + slot * St::kDoubles;
    // A x (all rows) and K x (rows < M): both need x only.
    {
      double acc = 0.0, acu = 0.0;
#pragma unroll 4
      for (int j = part; j < N; j += P) {
        const double xj = x[j];
        acc += blk[St::oA + j * N + row] * xj;
        if (row < M) acu += blk[St::oK + j * M + row] * xj;
      }
      red[part * N + row] = acc;
      if (row < M) red2[part * N + row] = acu;
    }
    __syncthreads();
    if (tid < N) ax[tid] = total(red, tid);
    if (tid >= kThreads - M) {  // u = k + K x, on threads of the last warp
      const int a = tid - (kThreads - M);
      const double ui = blk[St::ok + a] + total(red2, a);
      u[a] = ui;
      __stcs(uo + (static_cast<size_t>(k) * M + a) * L_, ui);
    }
    __syncthreads();
    // f = c' - delta' o v' + A x + B u
    {
      double acc = 0.0;
      for (int a = part; a < M; a += P) acc += blk[St::oB + a * N + row] * u[a];
      red[part * N + row] = acc;
    }
    __syncthreads();
    if (tid < N) {
      const double d = blk[St::od + tid];
      sdi[tid] = rsqrt(d);
      f[tid] = sdi[tid] * (blk[St::oc + tid] - d * blk[St::ov + tid] + ax[tid] + total(red, tid));
    }
    __syncthreads();
    // x' = D^1/2 P D^-1/2 f,  y' = v' + W' f
    red2[part * N + row] = w_times_f(blk + St::oW);
    __syncthreads();
    if (tid < N) {
      const double t = total(red2, tid);
      const double xi = blk[St::od + tid] * sdi[tid] * t;
      x[tid] = xi;
      __stcs(xo + (static_cast<size_t>(k + 1) * N + tid) * L_, xi);
      __stcs(yo + (static_cast<size_t>(k + 1) * N + tid) * L_,
             blk[St::ov + tid] + sdi[tid] * (f[tid] - t));
    }
    __syncthreads();  // this slot is free for the fetch of stage k + 2, x' is visible
  }
}

template <int N, int M>
struct CtaPlan {
  static_assert(N % 16 == 0 && N <= 64, "state dimension must be a multiple of 16, at most 64");
  static_assert(M % 4 == 0, "control dimension must be a multiple of 4 (MMA k step)");
  static int64_t store_elems(int T) { return CtaSizes<N, M>::store(T); }
  static int64_t scratch_elems(int T) { return CtaSizes<N, M>::scratch(T); }
  template <bool SOLVE>
  static void backward(const FastArgs &a, cudaStream_t s) {
    auto kern = riccati_backward_cta<N, M, SOLVE>;
    constexpr int bytes = CtaSmem<N, M>::kBytes;
    if (bytes > 48 * 1024)
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    ProfScope ps(a.prof, "riccati_backward_cta", s);
    kern<<<static_cast<unsigned>(a.batch), kThreads, bytes, s>>>(a.pm, a.status, a.store,
                                                                 a.scratch, a.batch, a.ld,
                                                                 a.num_edges);
  }
  static void forward(const FastArgs &a, cudaStream_t s) {
    auto kern = rollout_forward_cta<N, M>;
    constexpr int bytes = 2 * RolloutStage<N, M>::kDoubles * int(sizeof(double));
    if (bytes > 48 * 1024)
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    ProfScope ps(a.prof, "rollout_forward_cta", s);
    kern<<<static_cast<unsigned>(a.batch), kThreads, bytes, s>>>(a.pm, a.out, a.store, a.scratch,
                                                                 a.batch, a.ld, a.num_edges);
  }
  static int factor(const FastArgs &a, cudaStream_t s) {
    backward<false>(a, s);
    return 1;
  }
  static int solve(const FastArgs &a, cudaStream_t s) {
    {
      ProfScope ps(a.prof, "affine_backward_cta", s);
      affine_backward_cta<N, M><<<static_cast<unsigned>(a.batch), kThreads, 0, s>>>(
          a.pm, a.store, a.scratch, a.batch, a.ld, a.num_edges);
    }
    forward(a, s);
    return 2;
  }
  static int factor_solve(const FastArgs &a, cudaStream_t s) {
    backward<true>(a, s);
    forward(a, s);
    return 2;
  }
};

template <int N, int M>
const FastPlan *make_cta_plan(const char *name) {
  using P = CtaPlan<N, M>;
  static const FastPlan plan{name,         N,           M,         &P::store_elems,
                             &P::scratch_elems, &P::factor, &P::solve, &P::factor_solve,
                             nullptr, true};
  return &plan;
}

}  // namespace

const FastPlan *select_cta_plan(int n, int m) {
  if (n == 64 && m == 24) return make_cta_plan<64, 24>("cta_dmma_n64_m24");
  if (n == 16 && m == 4) return make_cta_plan<16, 4>("cta_dmma_n16_m4");
  if (n == 32 && m == 8) return make_cta_plan<32, 8>("cta_dmma_n32_m8");
  return nullptr;
}

}  // namespace sipoc
