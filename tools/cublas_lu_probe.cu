// Measurement probe, NOT product code: the reference's GPU build factors and solves its Jacobians with
// cublasDgetrfBatched / cublasDgetrsBatched through the pointer-array API
// (/root/reference/source/xnet_linalg.F90:1847 LUDecompBatched_GPU, :1938 LUBksubBatched_GPU, called from
// xnet_jacobian_dense.F90:587-589,686-688).  This library times exactly those two calls on a batch of
// caller-supplied matrices (column-major n x n, as jac(msize,msize,nz)), CUDA events on the stream the handle uses,
// so that bench.py can put "the kernel to beat" next to xb_jacobian_decomp / xb_jacobian_bksub on the same matrices.
// Build: make -C tools   (nvcc -shared ... -lcublas)  -> tools/libxb_cublas_probe.so
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) { std::snprintf(g_err, sizeof g_err, "%s: %s", #x, cudaGetErrorString(e_)); return -3; } \
  } while (0)
#define CB(x)                                                                      \
  do {                                                                             \
    cublasStatus_t s_ = (x);                                                       \
    if (s_ != CUBLAS_STATUS_SUCCESS) { std::snprintf(g_err, sizeof g_err, "%s: cublas status %d", #x, (int)s_); return -4; } \
  } while (0)

static char g_err[256];

extern "C" {

const char *xbp_last_error(void) { return g_err; }

// a[batch][n*n] column-major matrices, b[batch][n] right-hand sides (host).  reps timed repetitions after one warm-up,
// each on a fresh copy of the matrices (the factorisation is in place).  Outputs: best-of-reps milliseconds of the
// getrf and getrs calls, the factors / pivots / solutions of the last repetition (any may be NULL).
int xbp_cublas_getrf_getrs(int device, int n, int batch, const double *a, const double *b, int reps, float *ms_getrf,
                           float *ms_getrs, double *lu_out, int *ipiv_out, int *info_out, double *x_out) {
  CK(cudaSetDevice(device));
  cublasHandle_t h;
  CB(cublasCreate(&h));
  const size_t nn = (size_t)n * n;
  double *dA0 = nullptr, *dA = nullptr, *dB0 = nullptr, *dB = nullptr;
  int *dP = nullptr, *dInfo = nullptr;
  double **dAp = nullptr, **dBp = nullptr;
  CK(cudaMalloc(&dA0, nn * batch * 8)); CK(cudaMalloc(&dA, nn * batch * 8));
  CK(cudaMalloc(&dB0, (size_t)n * batch * 8)); CK(cudaMalloc(&dB, (size_t)n * batch * 8));
  CK(cudaMalloc(&dP, (size_t)n * batch * 4)); CK(cudaMalloc(&dInfo, (size_t)batch * 4));
  CK(cudaMalloc(&dAp, (size_t)batch * sizeof(double *))); CK(cudaMalloc(&dBp, (size_t)batch * sizeof(double *)));
  CK(cudaMemcpy(dA0, a, nn * batch * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB0, b, (size_t)n * batch * 8, cudaMemcpyHostToDevice));
  std::vector<double *> ap(batch), bp(batch);
  for (int i = 0; i < batch; ++i) { ap[i] = dA + nn * i; bp[i] = dB + (size_t)n * i; }
  CK(cudaMemcpy(dAp, ap.data(), (size_t)batch * sizeof(double *), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBp, bp.data(), (size_t)batch * sizeof(double *), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1, e2;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
  float best_f = 1e30f, best_s = 1e30f;
  for (int r = 0; r <= reps; ++r) {
    CK(cudaMemcpy(dA, dA0, nn * batch * 8, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(dB, dB0, (size_t)n * batch * 8, cudaMemcpyDeviceToDevice));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, 0));
    CB(cublasDgetrfBatched(h, n, dAp, n, dP, dInfo, batch));
    CK(cudaEventRecord(e1, 0));
    int hinfo = 0;
    CB(cublasDgetrsBatched(h, CUBLAS_OP_N, n, 1, (const double *const *)dAp, n, dP, dBp, n, &hinfo, batch));
    CK(cudaEventRecord(e2, 0));
    CK(cudaEventSynchronize(e2));
    float f, s;
    CK(cudaEventElapsedTime(&f, e0, e1)); CK(cudaEventElapsedTime(&s, e1, e2));
    if (r > 0) { if (f < best_f) best_f = f; if (s < best_s) best_s = s; }
  }
  if (ms_getrf) *ms_getrf = best_f;
  if (ms_getrs) *ms_getrs = best_s;
  if (lu_out) CK(cudaMemcpy(lu_out, dA, nn * batch * 8, cudaMemcpyDeviceToHost));
  if (ipiv_out) CK(cudaMemcpy(ipiv_out, dP, (size_t)n * batch * 4, cudaMemcpyDeviceToHost));
  if (info_out) CK(cudaMemcpy(info_out, dInfo, (size_t)batch * 4, cudaMemcpyDeviceToHost));
  if (x_out) CK(cudaMemcpy(x_out, dB, (size_t)n * batch * 8, cudaMemcpyDeviceToHost));
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  cudaFree(dA0); cudaFree(dA); cudaFree(dB0); cudaFree(dB); cudaFree(dP); cudaFree(dInfo); cudaFree(dAp); cudaFree(dBp);
  cublasDestroy(h);
  return 0;
}

}  // extern "C"
