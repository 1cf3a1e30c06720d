// DMMA (mma.sync m8n8k4 f64) issue-rate probe on sm_100a: FMA/clk/SM vs resident warps and independent accumulators.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe.bin tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k(double *out, int iters, long long *cyc) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = 0.5 * i; }
  const double a = 1e-3 * (threadIdx.x & 7), b = 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int threads, double *d, long long *dc) {
  const int iters = 4096;
  k<ILP><<<148, threads>>>(d, iters, dc); cudaDeviceSynchronize();
  k<ILP><<<148, threads>>>(d, iters, dc); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  double fma = (double)iters * ILP * (threads / 32) * 256.0 / (double)c;   // 8*8*4 FMA per mma
  printf("warps/SM %2d ILP %2d: %7.2f FMA/clk/SM  (%.1f cycles per mma per warp)\n", threads / 32, ILP, fma, (double)c / ((double)iters * ILP));
}
int main() {
  double *d; long long *dc;
  cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&dc, 8);
  for (int threads : {32, 128, 256, 512}) { run<1>(threads, d, dc); run<2>(threads, d, dc); run<4>(threads, d, dc); run<8>(threads, d, dc); }
  cudaError_t e = cudaGetLastError(); printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
