// common.cuh -- shared device/host helpers for libgbops (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/gbops.h"

#ifndef __CUDA_ARCH__
#define GB_HOST 1
#endif

namespace gb {

constexpr int kNumSMsB200 = 148;

// launch bookkeeping -------------------------------------------------------------------------------
// Host state is per device and safe to touch from several host threads (autograd runs backward on its own thread; one
// process may drive all eight GPUs): the counter is atomic, device facts are cached per device, tuning knobs are atomics.
extern std::atomic<unsigned long long> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

inline int finish_launch() {
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

constexpr int kMaxDevices = 64;
int num_sms();  // of the CURRENT device (api.cu)

// Raise the dynamic shared-memory limit of `kernel` on the current device to the device maximum minus the kernel's static
// usage -- once per (kernel, device), always to the same value, so concurrent launches never shrink each other's limit --
// and check that `needed` bytes fit.  Returns a cudaError_t as int.
int raise_smem_limit(const void *kernel, size_t needed);
template <class K>
inline int raise_smem_limit(K kernel, size_t needed) {
  return raise_smem_limit(reinterpret_cast<const void *>(kernel), needed);
}

// tuning knobs (host side) ---------------------------------------------------------------------------
struct Tuning {
  std::atomic<int> fps_cluster{0};
  std::atomic<int> fps_threads{0};
  std::atomic<int> fps_defer{0};   // 1: store every pick to global memory inside the round loop
  std::atomic<int> fps_direct{0};  // 1: CTA-level stage before the cluster exchange even when every warp could push directly
  std::atomic<int> group_split{0};
  std::atomic<int> group_mode{0};  // flags: 1 plain stores, 2 generic kernel, 4 no sorted backward, 8 no single-row TMA kernel, 16 flattened forward split
  std::atomic<int> group_ch{0};         // forward: channels staged per CTA (multiple of the interleave), 0 = automatic
  std::atomic<int> group_target_kb{0};  // forward: output per CTA of the aligned partition (0 = 512)
  std::atomic<int> interp_mode{0};
  std::atomic<int> query_qpw{0};
  std::atomic<int> scatter_cc{0};
  std::atomic<int> scatter_nt{0};  // dense backward: log2 of the target slots per CTA (8..10), 0 = smallest that holds n
  std::atomic<int> scatter_mode{0};  // bit 0: never use the dense (thread-owned targets) backward; bit 2: targets in index order (no degree sort);
                         // bit 3: never use the warp-private backward; bit 4: use it for any number of tasks
  std::atomic<int> priv_vl{0};     // warp-private backward: positions per lane and load (1, 2, 4), 0 = automatic
  std::atomic<int> priv_dry{0};    // experiment: 1 = the warp-private backward only streams its inputs (wrong results)
  std::atomic<int> priv_split{0};  // warp-private backward: warps that share a task (1, 2, 4), 0 = automatic
  std::atomic<int> priv_rows{0};   // warp-private backward, nsample 8 / 16: 1 = channel planes share a row (S = nsample), 2 = several rows per unit (S = 32), 0 = automatic
  std::atomic<int> priv_cw{0};     // warp-private backward: channels per warp and plane (2, 4), 0 = automatic
  std::atomic<int> query_mode{0};  // 1: never use the cell grid, 2: always use it (when the shape allows)
  std::atomic<int> grid_cell_pct{0};  // cell edge as a percentage of the query reach (0 = default 50)
};
extern Tuning g_tuning;

// stream-ordered scratch from the device's default memory pool (scatter.cu)
cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t s);

// scatter.cu: atomic-free segmented scatter-add shared by the group and interpolate backward passes
bool seg_scatter_supported(int b, int c, int n, size_t entries, int div);
int seg_scatter_add(const float *src, size_t src_stride, const int *key, const float *weight, float *grad, int b, int c, int n,
                    size_t entries, int div, int overwrite, cudaStream_t s);

// scatter_private.cu: group backward with warp-private shared-memory accumulators (no sort, no atomics)
bool scatter_private_supported(int b, int c, int n, int npoints, int nsample, size_t src_stride, const float *src, const int *idx);
int scatter_private(const float *src, size_t src_stride, const int *idx, float *grad, int b, int c, int n, int npoints, int nsample,
                    int overwrite, cudaStream_t s);

// query.cu: three nearest neighbours through the uniform cell grid (exact: same indices and distances as the full scan)
bool three_nn_grid_worth(int b, int n, int m);
int three_nn_grid(const float *unknown, const float *known, float *dist2, int *idx, float *weight, int b, int n, int m, cudaStream_t s);

// squared distance exactly as nvcc contracts the reference's (a-b)*(a-b)+(c-d)*(c-d)+(e-f)*(e-f):
// FMUL on the y term, then FFMA x, then FFMA z (SASS of ball_query_gpu.cu / sampling_gpu.cu / interpolate_gpu.cu).
__device__ __forceinline__ float sqdist3(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// streaming (evict-first) 128-bit global store and no-allocate 128-bit load
__device__ __forceinline__ void st_cs_f4(float *p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cs_f1(float *p, float v) { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ float4 ld_nc_na_f4(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_nc_i4(const int *p) {
  int4 v;
  asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// ---- mbarrier / bulk-copy (TMA engine) / cluster PTX wrappers ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// acquire at cluster scope: data written by a peer CTA's st.async is visible after this wait
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait_cluster(bar, parity)) {
  }
}

// 1-D bulk copy global -> shared through the TMA engine (UBLKCP); dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 1-D bulk copy shared -> global (UBLKCP store side) + group bookkeeping
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}
// map a local shared address to the same offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// asynchronous remote store + transaction-count completion on the REMOTE mbarrier (DSMEM push)
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr), "r"(a),
               "r"(b), "r"(c), "r"(d), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t remote_addr, uint32_t a, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(a), "r"(remote_bar)
               : "memory");
}

}  // namespace gb
