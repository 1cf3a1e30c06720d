// FP64 FMA issue-rate probe: throughput per SM as a function of resident warps and independent chains.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, long long *cyc) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double m = 0.999999, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int threads, double *d, long long *dc) {
  const int iters = 4096;
  k<ILP><<<148, threads>>>(d, iters, dc);
  cudaDeviceSynchronize();
  k<ILP><<<148, threads>>>(d, iters, dc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  double fma_per_clk_sm = (double)iters * ILP * threads / (double)c;
  printf("warps/SM %2d ILP %2d: %6.2f FMA/clk/SM  (%.1f cycles per warp-DFMA per warp)\n", threads / 32, ILP, fma_per_clk_sm,
         (double)c / ((double)iters * ILP));
}
int main() {
  double *d; long long *dc;
  cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&dc, 8);
  for (int threads : {32, 128, 256, 512, 1024}) {
    run<1>(threads, d, dc); run<2>(threads, d, dc); run<4>(threads, d, dc); run<8>(threads, d, dc); run<20>(threads, d, dc);
  }
  return 0;
}
