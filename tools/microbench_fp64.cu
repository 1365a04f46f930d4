// FP64 issue / latency microbenchmark for the B200 SM: DFMA and DMMA.8x8x4 dependent-chain
// latency, and DMMA throughput against the number of independent chains per warp and
// warps per SM.  The CTA-per-problem Riccati kernels are sized from these numbers.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/microbench_fp64 tools/microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void dmma_kernel(double *out, long long *cycles, int iters, double a, double b) {
  double acc[CHAINS][2];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c][0] = acc[c][1] = threadIdx.x + c;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) dmma(acc[c], a, b);
  }
  const long long t1 = clock64();
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CHAINS>
__global__ void dfma_kernel(double *out, long long *cycles, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x + c;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  const long long t1 = clock64();
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void rsqrt_kernel(double *out, long long *cycles, int iters, double a) {
  double x = a + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) x = rsqrt(x) + a;
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <class F>
static void run(const char *name, int chains, int warps, int iters, double ops_per_iter_per_warp, F launch) {
  double *out;
  long long *cyc, h = 0;
  cudaMalloc(&out, sizeof(double) * 1024 * 1024);
  cudaMalloc(&cyc, sizeof(long long));
  launch(out, cyc);
  launch(out, cyc);
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double per_iter = double(h) / iters;
  printf("%-6s chains/warp %2d  warps/SM %2d : %8.1f cycles/iter  %7.2f cycles per op per warp, "
         "%7.2f SM-cycles per op\n",
         name, chains, warps, per_iter, per_iter / ops_per_iter_per_warp,
         per_iter / (ops_per_iter_per_warp * warps));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  const int iters = 4096;
  const int warp_counts[] = {1, 4, 8, 16};
  for (int w : warp_counts) {
#define DM(C)                                                                          \
  run("DMMA", C, w, iters, C, [&](double *o, long long *c) {                           \
    dmma_kernel<C><<<1, 32 * w>>>(o, c, iters, 1.0000001, 0.9999999);                  \
  });
    DM(1) DM(2) DM(4) DM(8) DM(16)
#define DF(C)                                                                          \
  run("DFMA", C, w, iters, C, [&](double *o, long long *c) {                           \
    dfma_kernel<C><<<1, 32 * w>>>(o, c, iters, 1.0000001, 0.9999999);                  \
  });
    DF(1) DF(2) DF(4) DF(8)
  }
  run("RSQRT", 1, 1, iters, 1, [&](double *o, long long *c) { rsqrt_kernel<<<1, 32>>>(o, c, iters, 1.5); });
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
