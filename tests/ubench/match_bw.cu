// match_bw.cu -- cost of match.any.sync.b32 (dedupe of the targets of a warp step) on sm_100a: dependent chain (latency) and
// 1..16 warps per SM issuing independent matches (throughput).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o match_bw match_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int ITERS = 4096;
__global__ void k_chain(long long *out, int seed) {
  unsigned v = seed + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) v = __match_any_sync(0xffffffffu, (v + i) & 1023u) + threadIdx.x;
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0), out[1] = v;
}
__global__ void k_tput(long long *out, int seed) {
  unsigned v[4];
  for (int j = 0; j < 4; ++j) v[j] = seed * (j + 1) + threadIdx.x * 2654435761u;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc += __match_any_sync(0xffffffffu, (v[j] >> 7) & 1023u);
      v[j] = v[j] * 1664525u + 1013904223u;
    }
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[2 * (threadIdx.x >> 5)] = (t1 - t0), out[2 * (threadIdx.x >> 5) + 1] = acc;
}
int main() {
  long long *d, h[64];
  cudaMalloc(&d, sizeof(h));
  k_chain<<<1, 32>>>(d, 1);
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("match.any dependent chain: %.1f cycles\n", (double)h[0] / ITERS);
  for (int w = 1; w <= 32; w *= 2) {
    k_tput<<<1, 32 * w>>>(d, 3);
    cudaMemcpy(h, d, sizeof(long long) * 2 * w, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < w; ++i) mx = h[2 * i] > mx ? h[2 * i] : mx;
    printf("%2d warps on one SM: %.2f cycles per match per warp, %.2f cycles per match per SM\n", w, (double)mx / (4.0 * ITERS),
           (double)mx / (4.0 * ITERS * w));
  }
  return 0;
}
