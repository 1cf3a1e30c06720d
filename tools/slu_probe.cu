// Latency probe for the pivot loop of the structured LU (tools, not product): one CTA per SM, NW warps, each iteration =
// named barrier, U loads (table -> address -> value), 16 shared loads, 8 DFMA, 8 predicated stores.
// usage: slu_probe.bin   (prints cycles per iteration for several variants)
#include <cstdio>
#include <cuda_runtime.h>
extern __shared__ double sm[];
template <int MODE>
__global__ void probe(long long *out, int iters, int stride, int nw) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < 6000; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  unsigned *tab = reinterpret_cast<unsigned *>(sm + 6000);
  for (int i = tid; i < 512; i += blockDim.x) tab[i] = (unsigned)((i * stride) % 4000);
  __syncthreads();
  if (w >= nw) return;
  const unsigned sS = (unsigned)__cvta_generic_to_shared(sm);
  long long t0 = clock64();
  double rk = 1.0000001;
  for (int k = 0; k < iters; ++k) {
    asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
    if (MODE == 0) continue;                       // barrier only
    const unsigned tj = tab[(k + lane) & 511];     // table lookup
    const int cbj = (int)tj;
    const double u = sm[cbj + (k & 63)] * rk;      // dependent load + multiply
    if (MODE == 1) { if (__double2hiint(u) == 0x7ff12345) rk = 0; continue; }
    const unsigned lp = sS + 8u * (unsigned)(4000 + (k & 127) + 8 * w);
    const unsigned ap = sS + 8u * (unsigned)(cbj + (k & 63) + 1 + 8 * w);
    double l[8], a[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(l[t]) : "r"(lp + 8 * t));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[t]) : "r"(ap + 8 * t));
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] = fma(-l[t], u, a[t]);
    if (MODE == 2) { double s = 0; for (int t = 0; t < 8; ++t) s += a[t]; if (__double2hiint(s) == 0x7ff12345) rk = 0; continue; }
#pragma unroll
    for (int t = 0; t < 8; ++t) if (lane < 24) asm volatile("st.shared.f64 [%0], %1;" ::"r"(ap + 8 * t), "d"(a[t] * 1e-3));
    if (MODE == 4 && tid == 0) {                   // reciprocal of the next pivot by one thread
      double x = a[0], r;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
      double e = fma(-x, r, 1.0); e = fma(e, e, e); sm[5000 + (k & 255)] = fma(r, e, r);
    }
  }
  long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  if (rk == 0) out[0] = 0;
}
template <int MODE>
void run(const char *name, int nw, int stride, int grid) {
  long long *d; cudaMalloc(&d, 8 * grid);
  const int iters = 20000;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  probe<MODE><<<grid, 128, 70000>>>(d, iters, stride, nw);
  probe<MODE><<<grid, 128, 70000>>>(d, iters, stride, nw);
  long long h[1024]; cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
  printf("%-34s warps %d stride %3d ctas/SM %d: %7.1f cycles / iteration  (%s)\n", name, nw, stride, grid / 148, (double)h[0] / iters, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  for (int g : {148, 444}) {
    for (int nw : {1, 4}) {
      run<0>("barrier only", nw, 17, g);
      run<1>("+ table -> u load -> multiply", nw, 17, g);
      run<2>("+ 16 loads, 8 DFMA", nw, 17, g);
      run<3>("+ 8 stores (stride 17 columns)", nw, 17, g);
      run<3>("+ 8 stores (stride 16: conflicts)", nw, 16, g);
      run<4>("+ reciprocal by thread 0", nw, 17, g);
    }
  }
  return 0;
}
