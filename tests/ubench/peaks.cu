// peaks.cu -- measured arithmetic peaks the non-HBM kernels are reported against (bench.py `kernel_notes`):
// dependent-chain-free DFMA and FFMA loops, one CTA of 1024 threads per SM slot, 8 independent accumulators per thread.
// Built by __graft_entry__.build() into tests/ubench/libgb_peaks.so; C ABI: returns lane-operations per second (FMA = 1).
#include <cuda_runtime.h>
#include <stdint.h>

template <typename T>
__global__ void __launch_bounds__(1024) fma_loop(T *out, int iters, T a, T b) {
  T x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (T)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = x[i] * a + b;  // contracted to one FMA each
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == (T)12345.678) out[0] = s;  // never true: keeps the loop alive
}

template <typename T>
static double run(int iters) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  T *out = nullptr;
  if (cudaMalloc(&out, sizeof(T)) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int grid = sms * 2;
  fma_loop<T><<<grid, 1024>>>(out, 64, (T)1.0000001, (T)1e-9);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    fma_loop<T><<<grid, 1024>>>(out, iters, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) return -1.0;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * 1024.0 * 8.0 * (double)iters;
    if (ms > 0.f && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

extern "C" __attribute__((visibility("default"))) double ub_dfma_per_s(int iters) { return run<double>(iters); }
extern "C" __attribute__((visibility("default"))) double ub_ffma_per_s(int iters) { return run<float>(iters); }
