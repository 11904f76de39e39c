// fps.cu -- furthest point sampling as a thread-block-cluster kernel (sm_100a).
//
// Replaces furthest_point_sampling_kernel<BS> of PointNet/_ext_src/src/sampling_gpu.cu:74-178 (variant A) and of
// pointnet2_batch/src/sampling_gpu.cu:73-181 (variant B).  The reference runs ONE block per scene and re-reads every
// point and its running distance from global memory in each of the m-1 dependent rounds, with a 9-10 barrier shared
// memory tree per round.
//
// Here one CLUSTER of C CTAs owns a scene.  Each thread keeps P points (x, y, z, running min distance) in REGISTERS for
// the whole kernel, so a round touches no global memory at all:
//   1. every thread updates its P points against the last pick and keeps its best (distance, slot);
//   2. warp argmax with two redux.sync instructions on integer keys (distance bits, then tie key);
//   3. DIRECT mode (C > 1, at most 32 warps in the cluster): every warp PUSHES its own winner into the shared memory of
//      all C CTAs with st.async, which completes a transaction count on the receiver's mbarrier; every warp waits on its
//      CTA's mbarrier and reduces the <= 32 candidates, one per lane.  One DSMEM hop per round, no cluster barrier, no
//      __syncthreads, no CTA-level reduction.
//   3'. otherwise: warp winners -> shared memory -> one __syncthreads -> every warp reduces them redundantly, and (C > 1)
//      the CTA winner is pushed to all CTAs the same way.
// The picks are parked in shared memory during the rounds (the loop is latency-critical: a store or an index conversion
// inside it is paid m - 1 times) and written out once at the end, optionally with their coordinates (gb_fps_xyz).
// Segmented mode (gb_fps_segments): one CTA per segment of a packed point array, sizes from a device-side table.
//
// Bit-exact tie order.  The reference's result is determined by its block size BS = opt_n_threads(n): the per-thread
// strided scan keeps the first strict maximum and the shared-memory tree keeps the lower slot on ties, i.e. among equal
// distances the winner minimises (bitreverse_log2(BS)(k mod BS), k div BS).  That pair is packed in one 32-bit "tie
// key" and reduced with redux.min, so any decomposition (P, T, C) reproduces the reference's pick.
#include "common.cuh"

namespace gb {

constexpr int kKeyNone = (int)0xBF800000;  // bits of -1.0f as a signed int: below every valid (>= +0) distance

__device__ __forceinline__ uint32_t fps_tiekey(uint32_t k, int L) {
  const uint32_t low = L ? (__brev(k) >> (32 - L)) : 0u;  // bit reversal of (k mod BS) over L bits
  return (low << 22) | (k >> L);
}
__device__ __forceinline__ uint32_t fps_tiekey_inv(uint32_t tk, int L) {
  const uint32_t low = L ? (__brev(tk >> 22) >> (32 - L)) : 0u;
  return ((tk & 0x3FFFFFu) << L) | low;
}

struct FpsShared {
  uint32_t wc[2][32][8];  // per-warp candidates, double buffered by round parity: key, tiekey, x, y, z
  uint32_t cc[2][32][8];  // candidates received from the cluster: slot = sender rank (per-CTA winners), or sender rank * W + warp
                          // in direct mode (every warp's winner, at most 32 of them)
  uint64_t full[2];       // mbarriers: 20 bytes of st.async payload per candidate per round
};

template <int P, int T>
__global__ void __launch_bounds__(T, 1) fps_cluster_kernel(const float *__restrict__ xyz, float *__restrict__ temp,
                                                           int *__restrict__ idxs, int n, int m, int variant, int L, int direct,
                                                           float *__restrict__ new_xyz, int defer, const int4 *__restrict__ seg) {
  extern __shared__ float s_pts[];  // [3][P*T] copy of this CTA's coordinates (winner lookup without dynamic register indexing)
  __shared__ __align__(16) FpsShared sh;

  const uint32_t C = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int scene = blockIdx.x / C;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int W = T / 32;

  if (seg) {
    // segmented mode (one CTA per segment, C = 1): segments of a packed point array with their own sizes -- the per-object
    // FPS loop of ObjectBalanceSampling (TrainModel/modules.py:186-213) in one launch.  seg = (first point, points,
    // samples, first output slot); the tie order follows the segment's own BS = opt_n_threads(points).
    const int4 sg = seg[scene];
    xyz += (size_t)sg.x * 3;
    n = sg.y, m = sg.z;
    idxs += sg.w;
    if (new_xyz) new_xyz += (size_t)sg.w * 3;
    temp = nullptr;
    if (m <= 0) return;
    if (n <= 0) {  // nothing to sample from: index 0, as the reference's untouched besti
      for (int j = tid; j < m; j += T) idxs[j] = 0;
      return;
    }
    L = min(31 - __clz(n), variant == GB_FPS_A ? 9 : 10);
  } else {
    xyz += (size_t)scene * n * 3;
    idxs += (size_t)scene * m;
    if (temp) temp += (size_t)scene * n;
    if (new_xyz) new_xyz += (size_t)scene * m * 3;  // optional: coordinates of the picks (what gather_operation would fetch)
  }

  if (tid == 0) {
    mbar_init(&sh.full[0], 1);
    mbar_init(&sh.full[1], 1);
    fence_mbar_init();
  }

  float px[P], py[P], pz[P], pt[P];
  const uint32_t g = rank * T + tid;
  const uint32_t stride = C * T;
  float *sx = s_pts, *sy = s_pts + P * T, *sz = s_pts + 2 * P * T;
  uint32_t *s_picks = reinterpret_cast<uint32_t *>(s_pts + 3 * P * T);  // [m] tie keys of the picks (defer mode)
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const uint32_t k = g + i * stride;
    float x = 0.f, y = 0.f, z = 0.f, t = -1.0f;
    if (k < (uint32_t)n) {
      x = __ldg(xyz + 3 * (size_t)k), y = __ldg(xyz + 3 * (size_t)k + 1), z = __ldg(xyz + 3 * (size_t)k + 2);
      t = temp ? temp[k] : 1e10f;
      if (variant == GB_FPS_A) {
        const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
        if ((double)mag <= 1e-3) t = -1.0f;  // skipped: never updated, never picked (sampling_gpu.cu:105-106)
      }
    }
    px[i] = x, py[i] = y, pz[i] = z, pt[i] = t;
    sx[i * T + tid] = x, sy[i * T + tid] = y, sz[i * T + tid] = z;
  }
  // picked when nothing is eligible (reference: besti stays 0)
  const float p0x = __ldg(xyz), p0y = __ldg(xyz + 1), p0z = __ldg(xyz + 2);
  float cx = p0x, cy = p0y, cz = p0z;
  if (rank == 0 && tid == 0) idxs[0] = 0;

  __syncthreads();
  if (C > 1) cluster_sync_all();  // peers' mbarriers are initialised before anyone pushes

  uint32_t phases = 0u;  // bit p = parity to wait for on full[p]
  uint32_t cc0 = 0, cc1 = 0, bar0 = 0, bar1 = 0;  // this lane's push targets in CTA `lane` (slot = my rank), per parity
  if (C > 1 && (warp == 0 || direct) && lane < (int)C) {
    const int slot = direct ? (int)rank * W + warp : (int)rank;
    cc0 = mapa_u32(smem_u32(&sh.cc[0][slot][0]), lane), cc1 = mapa_u32(smem_u32(&sh.cc[1][slot][0]), lane);
    bar0 = mapa_u32(smem_u32(&sh.full[0]), lane), bar1 = mapa_u32(smem_u32(&sh.full[1]), lane);
  }
  const uint32_t ncand = direct ? C * W : C;  // candidates every CTA receives per round

  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    if (C > 1 && tid == 0) mbar_arrive_expect_tx(&sh.full[par], ncand * 20u);

    // 1. register-resident update + per-thread argmax (first strict maximum == lowest k among this thread's points)
    int bkey = kKeyNone, bi = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float d = sqdist3(px[i] - cx, py[i] - cy, pz[i] - cz);
      const float t = fminf(d, pt[i]);
      pt[i] = t;
      const int key = __float_as_int(t);
      if (key > bkey) bkey = key, bi = i;
    }
    // 2. warp argmax on (key desc, tiekey asc)
    const int wkey = __reduce_max_sync(0xffffffffu, bkey);
    const uint32_t tk = (bkey == wkey && bkey != kKeyNone) ? fps_tiekey(g + bi * stride, L) : 0xFFFFFFFFu;
    const uint32_t wtk = __reduce_min_sync(0xffffffffu, tk);
    const bool win = tk == wtk && bkey == wkey;  // unique lane unless the whole warp is ineligible (then all hold the same words)
    uint32_t ckey_u = (uint32_t)kKeyNone, ctk = 0xFFFFFFFFu, ux = 0, uy = 0, uz = 0;
    int key;
    uint32_t mtk, btk;
    int src;
    if (direct) {
      // 3'. DIRECT mode (C * W <= 32): every warp pushes its own winner to all C CTAs; no CTA-level stage at all (no shared
      // memory round trip, no __syncthreads, one reduction level less on the critical path of the round)
      if (win) {
        const int s = bi * T + tid;
        ux = __float_as_uint(sx[s]), uy = __float_as_uint(sy[s]), uz = __float_as_uint(sz[s]);
      }
      src = __ffs(__ballot_sync(0xffffffffu, win)) - 1;
      ux = __shfl_sync(0xffffffffu, ux, src), uy = __shfl_sync(0xffffffffu, uy, src), uz = __shfl_sync(0xffffffffu, uz, src);
      key = wkey, btk = wtk;
    } else {
      if (win) {
        uint32_t *w = sh.wc[par][warp];
        const int s = bi * T + tid;
        *reinterpret_cast<uint4 *>(w) = make_uint4((uint32_t)wkey, wtk, __float_as_uint(sx[s]), __float_as_uint(sy[s]));
        w[4] = __float_as_uint(sz[s]);
      }
      __syncthreads();
      // 3. every warp reduces the W warp candidates
      if (lane < W) {
        const uint4 v = *reinterpret_cast<const uint4 *>(sh.wc[par][lane]);
        ckey_u = v.x, ctk = v.y, ux = v.z, uy = v.w;
        uz = sh.wc[par][lane][4];
      }
      key = __reduce_max_sync(0xffffffffu, (int)ckey_u);
      mtk = ((int)ckey_u == key) ? ctk : 0xFFFFFFFFu;
      btk = __reduce_min_sync(0xffffffffu, mtk);
      src = __ffs(__ballot_sync(0xffffffffu, (int)ckey_u == key && ctk == btk)) - 1;
      ux = __shfl_sync(0xffffffffu, ux, src), uy = __shfl_sync(0xffffffffu, uy, src), uz = __shfl_sync(0xffffffffu, uz, src);
    }

    if (C > 1) {
      // 4. push the winner (of this CTA; of this warp in direct mode) to every CTA of the cluster (lane r -> CTA r), then
      // wait for all candidates
      if ((warp == 0 || direct) && lane < (int)C) {
        const uint32_t dst = par ? cc1 : cc0, bar = par ? bar1 : bar0;
        st_async_v4(dst, (uint32_t)key, btk, ux, uy, bar);
        st_async_b32(dst + 16, uz, bar);
      }
      mbar_wait_cluster(&sh.full[par], (phases >> par) & 1u);
      phases ^= 1u << par;
      ckey_u = (uint32_t)kKeyNone, ctk = 0xFFFFFFFFu;
      if (lane < (int)ncand) {
        const uint4 v = *reinterpret_cast<const uint4 *>(sh.cc[par][lane]);
        ckey_u = v.x, ctk = v.y, ux = v.z, uy = v.w;
        uz = sh.cc[par][lane][4];
      }
      key = __reduce_max_sync(0xffffffffu, (int)ckey_u);
      mtk = ((int)ckey_u == key) ? ctk : 0xFFFFFFFFu;
      btk = __reduce_min_sync(0xffffffffu, mtk);
      src = __ffs(__ballot_sync(0xffffffffu, (int)ckey_u == key && ctk == btk)) - 1;
      ux = __shfl_sync(0xffffffffu, ux, src), uy = __shfl_sync(0xffffffffu, uy, src), uz = __shfl_sync(0xffffffffu, uz, src);
    }
    if (key != kKeyNone) {
      cx = __uint_as_float(ux), cy = __uint_as_float(uy), cz = __uint_as_float(uz);
    } else {
      cx = p0x, cy = p0y, cz = p0z;
      btk = 0xFFFFFFFFu;
    }
    // defer mode: the round only parks the winner's tie key in shared memory; index conversion and the global stores
    // happen once, after the loop (the round loop is latency-critical: every instruction in it is paid m - 1 times)
    if (rank == 0 && tid == 0) {
      if (defer) s_picks[j] = btk;
      else idxs[j] = btk == 0xFFFFFFFFu ? 0 : (int)fps_tiekey_inv(btk, L);
    }
  }
  if (defer && rank == 0) {
    __syncthreads();
    for (int j = 1 + tid; j < m; j += T) {
      const uint32_t tk = s_picks[j];
      idxs[j] = tk == 0xFFFFFFFFu ? 0 : (int)fps_tiekey_inv(tk, L);
    }
  }

  // optional epilogue: coordinates of the picks, gathered here rather than stored round by round (three predicated
  // stores inside the round loop cost 5 % of the kernel even when unused: 1534 vs 1453 us for 20000 -> 2048)
  if (new_xyz && rank == 0) {
    __syncthreads();  // thread 0's index stores are visible to the CTA
    for (int j = tid; j < m; j += T) {
      const size_t k = (size_t)idxs[j];
      new_xyz[3 * j] = __ldg(xyz + 3 * k), new_xyz[3 * j + 1] = __ldg(xyz + 3 * k + 1), new_xyz[3 * j + 2] = __ldg(xyz + 3 * k + 2);
    }
  }
  if (temp) {
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const uint32_t k = g + i * stride;
      if (k < (uint32_t)n && pt[i] >= 0.f) temp[k] = pt[i];
    }
  }
  if (C > 1) cluster_sync_all();  // nobody leaves while a peer may still push into it
}

// Fallback for scenes too large to keep in the registers of one cluster: one CTA per scene, running distances in
// global memory (temp must be provided by the launcher), same key reduction.
template <int T>
__global__ void __launch_bounds__(T, 1) fps_global_kernel(const float *__restrict__ xyz, float *__restrict__ temp,
                                                          int *__restrict__ idxs, int n, int m, int variant, int L,
                                                          float *__restrict__ new_xyz) {
  __shared__ __align__(16) uint32_t wc[2][32][8];
  const int scene = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int W = T / 32;
  xyz += (size_t)scene * n * 3;
  idxs += (size_t)scene * m;
  temp += (size_t)scene * n;
  const float p0x = __ldg(xyz), p0y = __ldg(xyz + 1), p0z = __ldg(xyz + 2);
  float cx = p0x, cy = p0y, cz = p0z;
  if (new_xyz) new_xyz += (size_t)scene * m * 3;
  if (tid == 0) idxs[0] = 0;
  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    int bkey = kKeyNone;
    uint32_t bk = 0;
    for (uint32_t k = tid; k < (uint32_t)n; k += T) {
      const float x = __ldg(xyz + 3 * (size_t)k), y = __ldg(xyz + 3 * (size_t)k + 1), z = __ldg(xyz + 3 * (size_t)k + 2);
      if (variant == GB_FPS_A) {
        const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
        if ((double)mag <= 1e-3) continue;
      }
      const float t = fminf(sqdist3(x - cx, y - cy, z - cz), temp[k]);
      temp[k] = t;
      const int key = __float_as_int(t);
      if (key > bkey) bkey = key, bk = k;
    }
    const int wkey = __reduce_max_sync(0xffffffffu, bkey);
    const uint32_t tk = (bkey == wkey && bkey != kKeyNone) ? fps_tiekey(bk, L) : 0xFFFFFFFFu;
    const uint32_t wtk = __reduce_min_sync(0xffffffffu, tk);
    if (tk == wtk && bkey == wkey) {
      wc[par][warp][0] = (uint32_t)wkey;
      wc[par][warp][1] = wtk;
    }
    __syncthreads();
    uint32_t ckey_u = (uint32_t)kKeyNone, ctk = 0xFFFFFFFFu;
    if (lane < W) ckey_u = wc[par][lane][0], ctk = wc[par][lane][1];
    const int key = __reduce_max_sync(0xffffffffu, (int)ckey_u);
    const uint32_t btk = __reduce_min_sync(0xffffffffu, ((int)ckey_u == key) ? ctk : 0xFFFFFFFFu);
    int pick = 0;
    if (key != kKeyNone) pick = (int)fps_tiekey_inv(btk, L);
    cx = __ldg(xyz + 3 * (size_t)pick), cy = __ldg(xyz + 3 * (size_t)pick + 1), cz = __ldg(xyz + 3 * (size_t)pick + 2);
    if (tid == 0) idxs[j] = pick;
  }
  if (new_xyz) {
    __syncthreads();
    for (int j = tid; j < m; j += T) {
      const size_t k = (size_t)idxs[j];
      new_xyz[3 * j] = __ldg(xyz + 3 * k), new_xyz[3 * j + 1] = __ldg(xyz + 3 * k + 1), new_xyz[3 * j + 2] = __ldg(xyz + 3 * k + 2);
    }
  }
}

__global__ void fill_kernel(float *p, size_t n, float v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

constexpr int kFpsRetrySmallerCluster = -1001;  // internal: cluster shape not schedulable, try half the size

template <int P, int T>
static int launch_fps(const float *xyz, float *temp, int *idx, float *new_xyz, int b, int n, int m, int variant, int L, int C,
                      cudaStream_t s, const int4 *seg = nullptr) {
  auto kern = fps_cluster_kernel<P, T>;
  // defer mode parks the picks in shared memory (4 bytes each) when they fit beside the coordinate copy
  const int defer = ((size_t)3 * P * T * sizeof(float) + (size_t)m * 4 <= 200u * 1024u && g_tuning.fps_defer != 1) ? 1 : 0;
  const size_t dyn = (size_t)3 * P * T * sizeof(float) + (defer ? (size_t)m * 4 : 0);
  if (int rc_ = raise_smem_limit(kern, dyn)) return rc_;
  cudaError_t e = cudaSuccess;
  if (C > 8) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(b * C));
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)C;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (C > 1) {
    // can the device co-schedule a cluster of this size with this much shared memory?  (C = 16 is non-portable)
    int nclusters = 0;
    e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    if (e != cudaSuccess || nclusters < 1) {
      (void)cudaGetLastError();
      return kFpsRetrySmallerCluster;
    }
  }
  // direct mode: every warp's winner goes straight to all CTAs when they fit one lane each (g_tuning.fps_direct: 1 = never)
  const int direct = (C > 1 && C * (T / 32) <= 32 && g_tuning.fps_direct != 1) ? 1 : 0;
  e = cudaLaunchKernelEx(&cfg, kern, xyz, temp, idx, n, m, variant, L, direct, new_xyz, defer, seg);
  count_launch();
  return (int)e;
}

static int floor_log2(int v) {
  int l = 0;
  while ((1 << (l + 1)) <= v) ++l;
  return l;
}

}  // namespace gb

using namespace gb;

static int fps_impl(const float *xyz, float *temp, int *idx, float *new_xyz, int b, int n, int m, int variant, gb_stream_t stream,
                    int max_cluster = 0) {
  if (b < 0 || n < 0 || m < 0 || (variant != GB_FPS_A && variant != GB_FPS_B)) return (int)cudaErrorInvalidValue;
  if (b == 0 || m == 0) return 0;  // nothing to do (empty tensors have null data pointers)
  if (n == 0 || !xyz || !idx) return (int)cudaErrorInvalidValue;  // samples of an empty cloud are undefined (the reference reads out of bounds)
  cudaStream_t s = (cudaStream_t)stream;
  // BS = opt_n_threads(n): cuda_utils.h:21-27 (cap 512) / pointnet2_batch/src/cuda_utils.h:10-14 (cap 1024).
  // floor(log2 n) computed in integers (the reference's double log() agrees for every n < 2^31 that is not within
  // 1 ulp of a power of two from below -- checked around every power of two <= 2^22 in tests/test_cabi_and_host.py::test_block_size_formula_matches_the_reference_expression).
  const int cap = variant == GB_FPS_A ? 512 : 1024;
  int bs = 1 << floor_log2(n);
  if (bs > cap) bs = cap;
  const int L = floor_log2(bs);

  // ---- configuration: cluster size C, threads T, points per thread P, with C*T >= BS (tie-key alignment) ----
  const int sms = num_sms();
  int C = g_tuning.fps_cluster;
  if (C != 1 && C != 2 && C != 4 && C != 8 && C != 16) C = 0;
  if (C == 0) {
    // as many CTAs per scene as the chip holds at ONE CTA per SM (co-resident CTAs would share the issue slots of the
    // register sweep), but not fewer than ~1024 points per CTA (below that the DSMEM hop costs more than the shorter
    // sweep saves).  Measured on B200, 32 scenes of 20000 points: C=4 x 256 threads 0.75 us/round, C=8 x 1024 3.3 us.
    // (four CTAs of 128 threads pay from 512 points each: 2048 -> 1024 on 4-8 scenes 408 us against 434 us with two of 256)
    C = 8;
    while (C > 1 && ((long)b * C > (long)sms || n / C < (C == 4 ? 512 : 1024))) C >>= 1;
    // background sampling (the caller has the whole step to hide the rounds' latency): fewer, fuller CTAs -- each takes a
    // whole SM's registers, so the SMs they do not use are entirely free for the kernels they run beside
    if (max_cluster > 0 && C > max_cluster) C = max_cluster;
    // capacity: the largest per-CTA register tile is 256 threads x 40 points (= 512 x 20 = 1024 x 10)
    while (C < 16 && (long)C * 256 * 40 < n) C <<= 1;
  }
  for (; C >= 1; C >>= 1) {
    int T = g_tuning.fps_threads;
    if (T == 128 && (C * 4 > 32 || (n + C - 1) / C > 128 * 40)) T = 0;  // 128 threads: direct mode only, <= 40 points per thread
    if (T != 128 && T != 256 && T != 512 && T != 1024) {
      // the fewest warps that hold the CTA's points in registers: the per-round barrier and the redundant per-warp
      // reduction grow with the warp count, the sweep itself is issue-bound whatever the split
      const int per_cta = (n + C - 1) / C;
      T = per_cta <= 256 * 40 ? 256 : (per_cta <= 512 * 20 ? 512 : 1024);
      // 128 threads where that keeps every warp's winner on its own lane of the exchange (C * W <= 32: no CTA-level stage):
      // eight CTAs per scene (small shards: 20000 -> 2048 on 4 scenes 1089 us against 1212 us), or four CTAs of few points
      if (g_tuning.fps_direct != 1 && ((C == 8 && per_cta <= 128 * 40) || (C == 4 && per_cta <= 1024))) T = 128;
    }
    int Ce = C;
    while (Ce * T < bs && T < 1024) T <<= 1;
    while (Ce * T < bs && Ce < 16) Ce <<= 1;
    const int per_thread = (int)(((long)n + (long)Ce * T - 1) / ((long)Ce * T));
    int rc = kFpsRetrySmallerCluster - 1;  // "no instantiation"
#define GB_FPS_CASE(PP, TT) \
  else if (T == TT && per_thread <= PP) rc = launch_fps<PP, TT>(xyz, temp, idx, new_xyz, b, n, m, variant, L, Ce, s);
    if (false) {}
    GB_FPS_CASE(1, 1024) GB_FPS_CASE(2, 1024) GB_FPS_CASE(3, 1024) GB_FPS_CASE(4, 1024) GB_FPS_CASE(5, 1024)
    GB_FPS_CASE(6, 1024) GB_FPS_CASE(8, 1024) GB_FPS_CASE(10, 1024)
    GB_FPS_CASE(1, 512) GB_FPS_CASE(2, 512) GB_FPS_CASE(3, 512) GB_FPS_CASE(4, 512) GB_FPS_CASE(5, 512) GB_FPS_CASE(6, 512)
    GB_FPS_CASE(8, 512) GB_FPS_CASE(10, 512) GB_FPS_CASE(12, 512) GB_FPS_CASE(16, 512) GB_FPS_CASE(20, 512)
    GB_FPS_CASE(1, 256) GB_FPS_CASE(2, 256) GB_FPS_CASE(4, 256) GB_FPS_CASE(6, 256) GB_FPS_CASE(8, 256) GB_FPS_CASE(10, 256)
    GB_FPS_CASE(12, 256) GB_FPS_CASE(16, 256) GB_FPS_CASE(20, 256) GB_FPS_CASE(24, 256) GB_FPS_CASE(32, 256) GB_FPS_CASE(40, 256)
    GB_FPS_CASE(4, 128) GB_FPS_CASE(8, 128) GB_FPS_CASE(10, 128) GB_FPS_CASE(12, 128) GB_FPS_CASE(16, 128) GB_FPS_CASE(20, 128)
    GB_FPS_CASE(24, 128) GB_FPS_CASE(32, 128) GB_FPS_CASE(40, 128)
#undef GB_FPS_CASE
    if (rc == kFpsRetrySmallerCluster) continue;     // this cluster size cannot be scheduled: halve it
    if (rc == kFpsRetrySmallerCluster - 1) break;    // scene does not fit in the registers of C CTAs
    return rc;
  }

  // does not fit in registers: global-memory fallback (needs a temp buffer)
  float *tmp = temp;
  if (!tmp) {
    cudaError_t e = cudaMallocAsync((void **)&tmp, (size_t)b * n * sizeof(float), s);
    if (e != cudaSuccess) return (int)e;
    fill_kernel<<<1024, 256, 0, s>>>(tmp, (size_t)b * n, 1e10f);
    count_launch();
  }
  fps_global_kernel<1024><<<b, 1024, 0, s>>>(xyz, tmp, idx, n, m, variant, L, new_xyz);
  count_launch();
  int err = finish_launch();
  if (!temp) cudaFreeAsync(tmp, s);
  return err;
}

extern "C" int gb_fps(const float *xyz, float *temp, int *idx, int b, int n, int m, int variant, gb_stream_t stream) {
  return fps_impl(xyz, temp, idx, nullptr, b, n, m, variant, stream);
}

/* FPS + the gather_operation every caller runs on its result (pointnet2_modules.py:151-158: new_xyz = the sampled
 * coordinates), SURVEY 8f-2: new_xyz [b, m, 3] is written by the same launch -- the kernel holds the coordinates of
 * every pick anyway (they are the centre of the next round). */
extern "C" int gb_fps_xyz(const float *xyz, float *temp, int *idx, float *new_xyz, int b, int n, int m, int variant, gb_stream_t stream) {
  if (!new_xyz && b > 0 && m > 0) return (int)cudaErrorInvalidValue;
  return fps_impl(xyz, temp, idx, new_xyz, b, n, m, variant, stream);
}

/* gb_fps_xyz with a footprint hint: at most max_cluster (1, 2, 4, 8; 0 = automatic) CTAs per scene, as far as a scene still
 * fits their registers.  The automatic choice minimises the latency of the m - 1 dependent rounds (one CTA per SM, 4 CTAs per
 * scene for 32 scenes of 20000 points); a caller that samples the NEXT step's clouds beside the current step's work
 * (pipeline.OpPipeline.run(prefetch=...)) prefers 2: the rounds get 40 % slower but only 64 SMs are touched, and the
 * register-heavy kernels of the step keep the other 84 to themselves (B200, 32 scenes: 11.3 -> 10.9 ms per step).  Same picks. */
extern "C" int gb_fps_xyz_hint(const float *xyz, float *temp, int *idx, float *new_xyz, int b, int n, int m, int variant, int max_cluster,
                               gb_stream_t stream) {
  if (max_cluster < 0 || max_cluster > 16) return (int)cudaErrorInvalidValue;
  return fps_impl(xyz, temp, idx, new_xyz, b, n, m, variant, stream, max_cluster);
}

/* Segmented FPS: nseg independent point sets of different sizes packed in one array, one launch -- the per-object loop of
 * ObjectBalanceSampling (TrainModel/modules.py:186-213: `furthest_point_sample(object_points.unsqueeze(0), k)` per object
 * of every scene).  seg [nseg,4] i32 on the DEVICE = (first point, points, samples, first output slot) per segment;
 * idx (and new_xyz, optional) are indexed by output slot; indices are segment-local, ties resolved as gb_fps resolves
 * them for a cloud of the segment's size.  max_n / max_m = the largest points / samples entry (host knowledge; a segment
 * of up to 10240 points is held in the registers of one CTA, larger ones are refused with cudaErrorInvalidValue). */
extern "C" int gb_fps_segments(const float *xyz, const int *seg, int *idx, float *new_xyz, int nseg, int max_n, int max_m, int variant,
                               gb_stream_t stream) {
  if (nseg < 0 || max_n < 0 || max_m < 0 || (variant != GB_FPS_A && variant != GB_FPS_B)) return (int)cudaErrorInvalidValue;
  if (nseg == 0 || max_m == 0) return 0;
  if (!xyz || !seg || !idx) return (int)cudaErrorInvalidValue;
  if (((uintptr_t)seg & 15u) != 0) return (int)cudaErrorInvalidValue;
  const int T = variant == GB_FPS_A ? 512 : 1024;  // one CTA per segment: T must be a multiple of every segment's BS
  const int per_thread = (max_n + T - 1) / T;
  cudaStream_t s = (cudaStream_t)stream;
  const int4 *sg = reinterpret_cast<const int4 *>(seg);
  int rc = (int)cudaErrorInvalidValue;
#define GB_FPS_SEG(PP, TT) \
  else if (T == TT && per_thread <= PP) rc = launch_fps<PP, TT>(xyz, nullptr, idx, new_xyz, nseg, max_n, max_m, variant, 0, 1, s, sg);
  if (false) {}
  GB_FPS_SEG(1, 512) GB_FPS_SEG(2, 512) GB_FPS_SEG(4, 512) GB_FPS_SEG(6, 512) GB_FPS_SEG(8, 512) GB_FPS_SEG(10, 512) GB_FPS_SEG(12, 512)
  GB_FPS_SEG(16, 512) GB_FPS_SEG(20, 512)
  GB_FPS_SEG(1, 1024) GB_FPS_SEG(2, 1024) GB_FPS_SEG(4, 1024) GB_FPS_SEG(6, 1024) GB_FPS_SEG(8, 1024) GB_FPS_SEG(10, 1024)
#undef GB_FPS_SEG
  if (rc == kFpsRetrySmallerCluster) rc = (int)cudaErrorInvalidValue;
  return rc;
}
