// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
// Host launcher for the REFERENCE's own CUDA kernel.  The kernel itself is compiled by
// oracle/build_ref.py straight from /root/reference/kernels/fused_residual_rmsnorm.cu into
// oracle/_ref/fused_residual_rmsnorm_sm100a.cubin (the reference source is never copied); this file
// only loads that cubin and launches it with the reference's launch configuration
// (src/models/fused_ops.rs:161-181): grid = n_rows, block = n_cols < 1024 ? 32 : 1024, kernel
// arguments (x, residual, weight, dst, n_cols, block_size, eps), dst = [normed | sum].
#include <cuda_runtime.h>
#include <stdio.h>

static cudaLibrary_t g_lib = nullptr;

extern "C" int ref_load(const char* cubin_path) {
  if (g_lib) return 0;
  cudaError_t e = cudaLibraryLoadFromFile(&g_lib, cubin_path, nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) { fprintf(stderr, "ref_load: %s\n", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

static int launch(const char* name, const void* x, const void* r, const void* w, void* dst, int rows, int cols, float eps) {
  if (!g_lib) return -2;
  cudaKernel_t k;
  cudaError_t e = cudaLibraryGetKernel(&k, g_lib, name);
  if (e != cudaSuccess) return (int)e;
  int bs = cols < 1024 ? 32 : 1024;
  void* args[] = {(void*)&x, (void*)&r, (void*)&w, (void*)&dst, (void*)&cols, (void*)&bs, (void*)&eps};
  e = cudaLaunchKernel((const void*)k, dim3(rows), dim3(bs), args, 0, 0);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaDeviceSynchronize();
}

// host-buffer entry points: dst is [2*rows*cols] elements
static int run_host(const char* name, const void* x, const void* r, const void* w, void* dst, int rows, int cols, float eps, int es) {
  size_t n = (size_t)rows * cols * es;
  void *dx, *dr, *dw, *dd;
  if (cudaMalloc(&dx, n) || cudaMalloc(&dr, n) || cudaMalloc(&dw, (size_t)cols * es) || cudaMalloc(&dd, 2 * n)) return -1;
  cudaMemcpy(dx, x, n, cudaMemcpyHostToDevice);
  cudaMemcpy(dr, r, n, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, w, (size_t)cols * es, cudaMemcpyHostToDevice);
  int e = launch(name, dx, dr, dw, dd, rows, cols, eps);
  cudaMemcpy(dst, dd, 2 * n, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(dr); cudaFree(dw); cudaFree(dd);
  return e;
}
extern "C" int ref_fused_residual_rmsnorm_bf16_host(const void* x, const void* r, const void* w, void* dst, int rows, int cols, float eps) {
  return run_host("fused_residual_rmsnorm_bf16", x, r, w, dst, rows, cols, eps, 2);
}
extern "C" int ref_fused_residual_rmsnorm_f32_host(const void* x, const void* r, const void* w, void* dst, int rows, int cols, float eps) {
  return run_host("fused_residual_rmsnorm_f32", x, r, w, dst, rows, cols, eps, 4);
}
