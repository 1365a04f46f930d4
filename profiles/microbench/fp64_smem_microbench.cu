// Micro-benchmarks that size the shape-specialised Riccati kernel on B200:
//   * DFMA throughput (the FP64 roofline denominator that MEASURED_PEAKS.json
//     does not contain),
//   * shared-memory broadcast loads (8 distinct 16-byte addresses, each read by
//     4 lanes) in the two possible lane arrangements,
//   * SHFL throughput, and the latency of the dependent chains in a Cholesky
//     step (rsqrt / sqrt / div in FP64).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_smem_microbench.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__);        \
      return 1;                                                                \
    }                                                                          \
  } while (0)

__global__ void dfma_kernel(double *out, int iters) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 0: lanes 0-7 read 8 distinct 16B chunks, lanes 8-15 the same chunks, ...
// mode 1: lanes 0-3 read chunk 0, lanes 4-7 chunk 1, ...
// mode 2: 32 distinct chunks (no broadcast) for comparison.
// mode 3: like mode 0 but 8-byte loads (LDS.64).
__global__ void lds_kernel(double *out, int iters, int mode) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int chunk;
  if (mode == 0 || mode == 3) chunk = lane & 7;
  else if (mode == 1) chunk = lane >> 2;
  else chunk = lane;
  // per-problem region stride of 302 doubles (2 * odd) as in the solver
  const double *base = sm + (mode == 2 ? chunk * 2 : chunk * 302);
  double s0 = 0, s1 = 0;
  const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(base));
  if (mode == 3) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr + 16 * k));
        s0 += v;
      }
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        double vx, vy;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(addr + 16 * k));
        s0 += vx;
        s1 += vy;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1;
}

__global__ void shfl_kernel(double *out, int iters) {
  double v = threadIdx.x;
  double s = 0;
  const int src = (threadIdx.x & 7) + 8;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      v = __shfl_sync(0xffffffffu, v + 1.0, src);
      s += v;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Single-thread dependent chains, timed with clock64.
__global__ void latency_kernel(double *out, long long *cycles) {
  double x = 1.2345 + threadIdx.x;
  long long t0, t1;
  const int R = 256;
  t0 = clock64();
  for (int i = 0; i < R; ++i) x = fma(x, 1.0000001, 1e-9);
  t1 = clock64();
  cycles[0] = (t1 - t0) / R;
  t0 = clock64();
  for (int i = 0; i < R; ++i) x = rsqrt(x) + 1.5;
  t1 = clock64();
  cycles[1] = (t1 - t0) / R;
  t0 = clock64();
  for (int i = 0; i < R; ++i) x = sqrt(x) + 1.5;
  t1 = clock64();
  cycles[2] = (t1 - t0) / R;
  t0 = clock64();
  for (int i = 0; i < R; ++i) x = 1.0 / x + 1.5;
  t1 = clock64();
  cycles[3] = (t1 - t0) / R;
  t0 = clock64();
  for (int i = 0; i < R; ++i) x = __shfl_sync(0xffffffffu, x, 3) + 1.0;
  t1 = clock64();
  cycles[4] = (t1 - t0) / R;
  out[threadIdx.x] = x;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double *out;
  CK(cudaMalloc(&out, sizeof(double) * sms * 1024 * 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms;

  {  // DFMA peak: 8 CTAs x 256 threads per SM, 16 independent chains each
    const int iters = 20000;
    dfma_kernel<<<sms * 8, 256>>>(out, 100);
    CK(cudaDeviceSynchronize());
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      dfma_kernel<<<sms * 8, 256>>>(out, iters);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double flops = 2.0 * 16 * iters * sms * 8 * 256;
      printf("DFMA: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz nominal)\n", ms,
             flops / ms * 1e-9, flops / 2 / (ms * 1e-3) / sms / 1.965e9);
    }
  }
  {
    const int iters = 4000;
    for (int mode = 0; mode < 4; ++mode) {
      lds_kernel<<<sms, 512>>>(out, 10, mode);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      lds_kernel<<<sms, 512>>>(out, iters, mode);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaGetLastError());
      const double instr = 16.0 * iters * 16;  // warp-instructions per SM (16 warps)
      printf("LDS mode %d: %.3f ms  %.2f clk per warp-instruction per SM (1.965 GHz)\n", mode,
             ms, ms * 1e-3 * 1.965e9 / instr);
    }
  }
  {
    const int iters = 4000;
    shfl_kernel<<<sms, 512>>>(out, 10);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    shfl_kernel<<<sms, 512>>>(out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double instr = 16.0 * iters * 16 * 2;  // a double shuffle = 2 SHFL
    printf("SHFL(double): %.3f ms  %.2f clk per 32-bit SHFL warp-instruction per SM\n", ms,
           ms * 1e-3 * 1.965e9 / instr);
  }
  {
    long long *cyc;
    CK(cudaMalloc(&cyc, 8 * sizeof(long long)));
    latency_kernel<<<1, 32>>>(out, cyc);
    CK(cudaDeviceSynchronize());
    long long h[5];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("latency (cycles): dfma %lld  rsqrt+add %lld  sqrt+add %lld  div+add %lld  "
           "shfl(double)+add %lld\n", h[0], h[1], h[2], h[3], h[4]);
  }
  return 0;
}
