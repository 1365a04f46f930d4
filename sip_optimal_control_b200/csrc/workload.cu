#include "workload.cuh"

namespace sipoc {

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ inline uint64_t counter_hash(uint64_t seed, uint64_t array, uint64_t problem,
                                        uint64_t element, uint64_t draw) {
  uint64_t h = mix64(seed ^ (array * 0xD6E8FEB86659FD93ull));
  h = mix64(h ^ problem);
  h = mix64(h ^ (element * 2 + draw));
  return h;
}

// (0, 1]
__device__ inline double uniform01(uint64_t bits) {
  return (static_cast<double>(bits >> 11) + 1.0) * (1.0 / 9007199254740992.0);
}

__device__ inline double normal01(uint64_t seed, uint64_t array, uint64_t problem,
                                  uint64_t element) {
  const double u1 = uniform01(counter_hash(seed, array, problem, element, 0));
  const double u2 = uniform01(counter_hash(seed, array, problem, element, 1));
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

enum Kind : int { kA = 0, kB, kM, kR, kQ, kq, kr, kc, kDelta };

// grid.x covers the batch (coalesced stores), grid.y strides over flat elements.
__global__ void __launch_bounds__(128)
generate_kernel(int kind, uint64_t seed, int64_t problem_offset, int n, int m, int64_t size,
                int64_t batch, int64_t ld, double *out) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= ld) return;
  const uint64_t gp = static_cast<uint64_t>(problem_offset + b);
  for (int64_t e = blockIdx.y; e < size; e += gridDim.y) {
    double v = 0.0;
    if (b < batch) {
      switch (kind) {
        case kA: {
          const int64_t within = e % (static_cast<int64_t>(n) * n);
          const int row = static_cast<int>(within % n), col = static_cast<int>(within / n);
          v = 0.05 * normal01(seed, kA, gp, e) + (row == col ? 1.0 : 0.0);
          break;
        }
        case kB:
          v = 0.1 * normal01(seed, kB, gp, e);
          break;
        case kM:
          v = 0.0;
          break;
        case kR:
        case kQ: {
          const int d = kind == kR ? m : n;
          const int64_t dd = static_cast<int64_t>(d) * d;
          const int64_t stage = e / dd, within = e % dd;
          const int row = static_cast<int>(within % d), col = static_cast<int>(within / d);
          const int lo = row < col ? row : col, hi = row < col ? col : row;
          double s = 0.0;
          for (int k = 0; k < d; ++k) {
            const double zl = normal01(seed, kind, gp, stage * dd + k + static_cast<int64_t>(lo) * d);
            const double zh = normal01(seed, kind, gp, stage * dd + k + static_cast<int64_t>(hi) * d);
            s += zl * zh;
          }
          v = s + (row == col ? (kind == kR ? 1.01 : 1e-3) : 0.0);
          break;
        }
        case kq:
        case kr:
        case kc:
          v = normal01(seed, kind, gp, e);
          break;
        default:  // kDelta
          v = 1e-3 + 1e-1 * uniform01(counter_hash(seed, kDelta, gp, e, 0));
          break;
      }
    }
    out[e * ld + b] = v;
  }
}

}  // namespace

int launch_generate_lqr_benchmark(uint64_t seed, int64_t problem_offset, int E, int n, int m,
                                  int64_t batch, int64_t ld, double *Q, double *M, double *R,
                                  double *q, double *r, double *A, double *B, double *c,
                                  double *delta, cudaStream_t stream) {
  struct Job {
    int kind;
    double *out;
    int64_t size;
  };
  const int64_t nn = static_cast<int64_t>(n) * n, nm = static_cast<int64_t>(n) * m,
                mm = static_cast<int64_t>(m) * m;
  const Job jobs[9] = {{kA, A, E * nn},       {kB, B, E * nm},       {kM, M, E * nm},
                       {kR, R, E * mm},       {kQ, Q, (E + 1) * nn}, {kq, q, (E + 1) * (int64_t)n},
                       {kr, r, E * (int64_t)m}, {kc, c, (E + 1) * (int64_t)n},
                       {kDelta, delta, (E + 1) * (int64_t)n}};
  int launches = 0;
  for (const Job &j : jobs) {
    if (j.size == 0) continue;
    dim3 grid(static_cast<unsigned>((ld + 127) / 128),
              static_cast<unsigned>(j.size < 16384 ? j.size : 16384));
    generate_kernel<<<grid, 128, 0, stream>>>(j.kind, seed, problem_offset, n, m, j.size, batch,
                                              ld, j.out);
    ++launches;
  }
  return launches;
}

}  // namespace sipoc
