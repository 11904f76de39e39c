// latency.cu -- micro-benchmarks (SM cycles via clock64) of the primitives the FPS round is built from.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o latency latency.cu ; run on a B200.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../../graspbalance_b200/csrc/common.cuh"
namespace gb { unsigned long long g_launch_count = 0; Tuning g_tuning; }
using namespace gb;

constexpr int ITERS = 2000;

__global__ void k_redux(long long *out, int seed) {
  int v = seed + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) v = __reduce_max_sync(0xffffffffu, v + i) ^ threadIdx.x;
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / ITERS, out[1] = v;
}
__global__ void k_shfl(long long *out, int seed) {
  int v = seed + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) v = __shfl_xor_sync(0xffffffffu, v + i, 1);
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / ITERS, out[1] = v;
}
__global__ void k_ballot(long long *out, int seed) {
  int v = seed + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) v = __ffs(__ballot_sync(0xffffffffu, (v + i) & 1)) + v;
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / ITERS, out[1] = v;
}
__global__ void k_sync(long long *out) {
  __shared__ int s[1024];
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
    s[threadIdx.x] = i;
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / ITERS, out[1] = s[5];
}
// smem write -> syncthreads -> every warp reads 32 values + 2 redux (the CTA stage of the FPS round)
__global__ void k_cta_stage(long long *out) {
  __shared__ int s[2][32];
  int v = threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x / 32;
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
    int w = __reduce_max_sync(0xffffffffu, v + i);
    int w2 = __reduce_min_sync(0xffffffffu, (v == w) ? lane : 99);
    if (lane == w2 || lane == 0) s[i & 1][warp] = w;
    __syncthreads();
    int c = lane < W ? s[i & 1][lane] : 0;
    int m = __reduce_max_sync(0xffffffffu, c);
    int m2 = __reduce_min_sync(0xffffffffu, c == m ? lane : 99);
    v = __shfl_sync(0xffffffffu, c, m2 & 31) + threadIdx.x;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / ITERS, out[1] = v;
}
// cluster exchange A: st.async push to all CTAs + local mbarrier wait (what fps.cu does)
__global__ void k_push(long long *out) {
  __shared__ __align__(16) uint32_t cc[2][16][8];
  __shared__ uint64_t full[2];
  const uint32_t C = cluster_nctarank(), rank = cluster_ctarank();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); fence_mbar_init(); }
  __syncthreads();
  cluster_sync_all();
  uint32_t phases = 0, v = threadIdx.x;
  uint32_t cc0 = mapa_u32(smem_u32(&cc[0][rank][0]), lane % C), cc1 = mapa_u32(smem_u32(&cc[1][rank][0]), lane % C);
  uint32_t b0 = mapa_u32(smem_u32(&full[0]), lane % C), b1 = mapa_u32(smem_u32(&full[1]), lane % C);
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
    const int par = i & 1;
    if (threadIdx.x == 0) mbar_arrive_expect_tx(&full[par], C * 20u);
    if (warp == 0 && lane < (int)C) {
      st_async_v4(par ? cc1 : cc0, v, v + 1, v + 2, v + 3, par ? b1 : b0);
      st_async_b32((par ? cc1 : cc0) + 16, v + 4, par ? b1 : b0);
    }
    mbar_wait_cluster(&full[par], (phases >> par) & 1u);
    phases ^= 1u << par;
    v = cc[par][lane % C][0] + 1;
  }
  long long t1 = clock64();
  cluster_sync_all();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / ITERS, out[1] = v;
}
// cluster exchange B: plain DSMEM stores + cluster barrier
__global__ void k_barrier(long long *out) {
  __shared__ __align__(16) uint32_t cc[2][16][8];
  const uint32_t C = cluster_nctarank(), rank = cluster_ctarank();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cluster_sync_all();
  uint32_t v = threadIdx.x;
  uint32_t cc0 = mapa_u32(smem_u32(&cc[0][rank][0]), lane % C), cc1 = mapa_u32(smem_u32(&cc[1][rank][0]), lane % C);
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
    const int par = i & 1;
    if (warp == 0 && lane < (int)C) {
      uint32_t a = par ? cc1 : cc0;
      asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3) : "memory");
      asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a + 16), "r"(v + 4) : "memory");
    }
    cluster_sync_all();
    v = cc[par][lane % C][0] + 1;
  }
  long long t1 = clock64();
  cluster_sync_all();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / ITERS, out[1] = v;
}
// remote DSMEM load latency (pointer chase through the peer's shared memory)
__global__ void k_dsmem_ld(long long *out) {
  __shared__ uint32_t chain[64];
  const uint32_t C = cluster_nctarank(), rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 64; i += blockDim.x) chain[i] = (i + 1) & 63;
  __syncthreads();
  cluster_sync_all();
  uint32_t peer = mapa_u32(smem_u32(chain), (rank + 1) % C), v = 0;
  long long t0 = clock64();
  if (threadIdx.x == 0)
    for (int i = 0; i < ITERS; ++i) asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(peer + v * 4) : "memory");
  long long t1 = clock64();
  cluster_sync_all();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / ITERS, out[1] = v;
}

template <typename K, typename... A>
static long long run_cluster(K kern, int C, int T, long long *d, A... a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C); cfg.blockDim = dim3(T);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (C > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, d, a...);
  if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return -1; }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("run failed: %s\n", cudaGetErrorString(e)); return -1; }
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  return h[0];
}

int main() {
  long long *d; cudaMalloc(&d, 64);
  long long h[2];
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SM clock attr %d kHz\n", clk);
  for (int rep = 0; rep < 2; ++rep) {
    k_redux<<<1, 32>>>(d, 1); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("redux.max chain        : %lld cyc\n", h[0]);
    k_shfl<<<1, 32>>>(d, 1); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("shfl chain             : %lld cyc\n", h[0]);
    k_ballot<<<1, 32>>>(d, 1); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("ballot+ffs chain       : %lld cyc\n", h[0]);
    for (int T : {128, 256, 512, 1024}) {
      k_sync<<<1, T>>>(d); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("STS+__syncthreads T=%4d: %lld cyc\n", T, h[0]);
      k_cta_stage<<<1, T>>>(d); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("CTA stage T=%4d       : %lld cyc\n", T, h[0]);
    }
    for (int C : {2, 4, 8, 16}) {
      printf("cluster C=%2d T=512: st.async push+mbar wait %lld cyc | st.shared::cluster+cluster barrier %lld cyc | remote ld %lld cyc\n", C,
             run_cluster(k_push, C, 512, d), run_cluster(k_barrier, C, 512, d), run_cluster(k_dsmem_ld, C, 64, d));
      printf("cluster C=%2d T=128: st.async push+mbar wait %lld cyc | st.shared::cluster+cluster barrier %lld cyc\n", C,
             run_cluster(k_push, C, 128, d), run_cluster(k_barrier, C, 128, d));
    }
  }
  // wall-clock check of the SM clock: 1e8 dependent FMAs
  return 0;
}
