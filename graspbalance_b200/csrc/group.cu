// group.cu -- grouping_operation / gather_operation, forward and backward.
//
// Replaces group_points_kernel / group_points_grad_kernel (PointNet/_ext_src/src/group_points_gpu.cu:17-101; one block per
// scene, strided 4-byte writes), group_points_kernel_fast / _grad_kernel_fast (pointnet2_batch/src/group_points_gpu.cu:9-70;
// one thread per element, idx re-read for every channel, 4-byte sector-wasting gathers from L2) and the gather kernels
// (sampling_gpu.cu:13-62, batch :8-63).
//
// Forward is the HBM-bound op of the pipeline: out[b,c,j,k] = points[b,c,idx[b,j,k]] writes 4*C*npoints*nsample bytes per
// scene and reads only 4*C*n.  Design:
//   * a CTA stages the source rows of a chunk of channels in SHARED MEMORY (up to ~200 KB), interleaved V channels per
//     point, so that the random gather is one LDS.(32*V) per index instead of V sector-sized L2 reads;
//   * each thread owns 4 consecutive output positions: one 128-bit load of idx (read once per channel CHUNK, not once
//     per channel), 4 LDS gathers per channel group, one coalesced 128-bit streaming store per channel;
//   * the (scene, chunk) x position work space is flattened and cut into equal contiguous ranges, one per resident CTA,
//     so all 148 SMs finish together whatever the shape.
// Backward is a scatter-add: coalesced 128-bit reads of grad_out and idx, red.global.add.f32 into the (L2-resident)
// gradient rows, with WARP-AGGREGATION of runs of equal indices first -- ball/cylinder query pads a neighbourhood with
// copies of its first hit, so sparse neighbourhoods collapse to one atomic per run.
#include "common.cuh"

namespace gb {

constexpr int kGroupThreads = 512;

template <int V> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int V> __device__ __forceinline__ typename VecT<V>::type vmake(const float *v);
template <> __device__ __forceinline__ float vmake<1>(const float *v) { return v[0]; }
template <> __device__ __forceinline__ float2 vmake<2>(const float *v) { return make_float2(v[0], v[1]); }
template <> __device__ __forceinline__ float4 vmake<4>(const float *v) { return make_float4(v[0], v[1], v[2], v[3]); }

template <int V> __device__ __forceinline__ float vget(const typename VecT<V>::type &a, int v);
template <> __device__ __forceinline__ float vget<1>(const float &a, int) { return a; }
template <> __device__ __forceinline__ float vget<2>(const float2 &a, int v) { return v == 0 ? a.x : a.y; }
template <> __device__ __forceinline__ float vget<4>(const float4 &a, int v) { return v == 0 ? a.x : (v == 1 ? a.y : (v == 2 ? a.z : a.w)); }

// points [b,c,n]; idx [b,per]; out [b,c,per]; per % 4 == 0.  CH = channels staged per fill (multiple of V),
// chunks = ceil(c / CH), per4 = per / 4, wpc = work (quads) per CTA.
template <int V>
__global__ void __launch_bounds__(kGroupThreads) group_fwd_kernel(const float *__restrict__ points, const int *__restrict__ idx,
                                                                 float *__restrict__ out, int c, int n, int per4, int CH, int chunks,
                                                                 long long total, long long wpc, int streaming) {
  extern __shared__ __align__(16) float s_rows[];  // [CH/V][n][V]
  using Vec = typename VecT<V>::type;
  Vec *srow = reinterpret_cast<Vec *>(s_rows);
  const int tid = threadIdx.x;
  long long w = (long long)blockIdx.x * wpc;
  const long long wend = min(total, w + wpc);
  const int G = CH / V;
  const size_t per = (size_t)per4 * 4;

  while (w < wend) {
    const long long pair = w / per4;
    const int q0 = (int)(w - pair * per4);
    const int q1 = (int)min((long long)per4, (long long)q0 + (wend - w));
    const int scene = (int)(pair / chunks), chunk = (int)(pair - (long long)scene * chunks);
    const int ch_base = chunk * CH;
    const int gcount = min(G, (c - ch_base + V - 1) / V);

    // ---- fill: rows of this chunk, V channels interleaved per point ----
    __syncthreads();
    for (int g = 0; g < gcount; ++g) {
      const float *src = points + ((size_t)scene * c + ch_base + g * V) * n;
      const int nv = min(V, c - (ch_base + g * V));
      for (int i = tid; i < n; i += kGroupThreads) {
        float v[V];
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = e < nv ? __ldg(src + (size_t)e * n + i) : 0.f;
        srow[(size_t)g * n + i] = vmake<V>(v);
      }
    }
    __syncthreads();

    // ---- sweep: 4 consecutive positions per thread ----
    const int *ip = idx + (size_t)scene * per;
    for (int q = q0 + tid; q < q1; q += kGroupThreads) {
      const int4 id = ld_nc_i4(ip + (size_t)q * 4);
      for (int g = 0; g < gcount; ++g) {
        const Vec *row = srow + (size_t)g * n;
        const Vec a0 = row[id.x], a1 = row[id.y], a2 = row[id.z], a3 = row[id.w];
        const int ch0 = ch_base + g * V;
#pragma unroll
        for (int e = 0; e < V; ++e) {
          if (ch0 + e < c) {
            float *dst = out + ((size_t)scene * c + ch0 + e) * per + (size_t)q * 4;
            const float4 o = make_float4(vget<V>(a0, e), vget<V>(a1, e), vget<V>(a2, e), vget<V>(a3, e));
            if (streaming) st_cs_f4(dst, o);
            else *reinterpret_cast<float4 *>(dst) = o;
          }
        }
      }
    }
    w += (q1 - q0);
  }
}

// generic fallback (any shape / alignment): one thread per output element, gathers straight from global/L2
__global__ void group_fwd_generic_kernel(const float *__restrict__ points, const int *__restrict__ idx, float *__restrict__ out, int c,
                                         int n, size_t per, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / per, pos = e - row * per;
    const size_t scene = row / c;
    out[e] = __ldg(points + row * n + __ldg(idx + scene * per + pos));
  }
}

// ---- backward: warp-aggregated scatter-add ----------------------------------------------------------------------
// Sum `val` over runs of adjacent lanes with equal `key`; returns true on the first lane of each run (which then holds
// the run total).  Runs are found with one ballot; the doubling steps never cross a run boundary.
__device__ __forceinline__ bool warp_run_reduce(int key, float &val) {
  const unsigned lane = lane_id();
  const int prev = __shfl_up_sync(0xffffffffu, key, 1);
  const bool head = (lane == 0) || (prev != key);
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  // lanes strictly after me up to the next head belong to my run
  const unsigned after = lane == 31 ? 0u : (heads >> (lane + 1));
  const int run = after ? (__ffs(after) - 1) : (31 - (int)lane);  // number of followers in my run
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float other = __shfl_down_sync(0xffffffffu, val, d);
    // my partial covers min(run+1, d) lanes; add the block starting d lanes away if it is still inside my run
    if ((int)d <= run) val += other;
  }
  return head;
}

// grad_out [b,c,per]; idx [b,per]; grad_points [b,c,n] (+=).  One thread per 4 consecutive positions; grid.y = rows (b*c).
__global__ void __launch_bounds__(256) group_bwd_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx,
                                                        float *__restrict__ grad_points, int c, int n, int per4) {
  const size_t row = blockIdx.y;
  const size_t scene = row / c;
  const size_t per = (size_t)per4 * 4;
  const float *g = grad_out + row * per;
  const int *ip = idx + scene * per;
  float *dst = grad_points + row * n;
  const int qbase = blockIdx.x * (256 * 4);
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int q = qbase + it * 256 + threadIdx.x;
    const bool ok = q < per4;  // warp-uniform except in the last warp; shuffles below are executed by all lanes
    int4 id = make_int4(-1, -2, -3, -4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) {
      id = ld_nc_i4(ip + (size_t)q * 4);
      v = ld_nc_na_f4(g + (size_t)q * 4);
    }
    const bool uniform = ok && id.x == id.y && id.y == id.z && id.z == id.w;
    // lanes whose quad is one repeated index (padding) join a run reduction; others get a unique negative key
    int key = uniform ? id.x : -1 - (int)lane_id();
    float val = uniform ? ((v.x + v.y) + (v.z + v.w)) : 0.f;
    const unsigned any_uniform = __ballot_sync(0xffffffffu, uniform);
    bool head = true;
    if (any_uniform) head = warp_run_reduce(key, val);
    if (uniform) {
      if (head) atomicAdd(dst + id.x, val);
    } else if (ok) {
      // within-quad combining of adjacent equal indices, then one red per distinct run
      float acc = v.x;
      if (id.y == id.x) acc += v.y; else { atomicAdd(dst + id.x, acc); acc = v.y; }
      if (id.z == id.y) acc += v.z; else { atomicAdd(dst + id.y, acc); acc = v.z; }
      if (id.w == id.z) acc += v.w; else { atomicAdd(dst + id.z, acc); acc = v.w; }
      atomicAdd(dst + id.w, acc);
    }
  }
}

// ---- backward without atomics: per-tile counting sort + shared-memory row accumulators -------------------------------
// The scatter-add grad[b,c,idx[b,e]] += g[b,c,e] is rewritten as a segmented sum.  Pass 1 (once per call, shared by all
// channels): every tile of kBwdTile consecutive positions is counting-sorted by target index in shared memory, giving
// skey (sorted targets) and tperm (the tile-local position of each sorted entry).  Pass 2: a CTA owns CC channel rows of
// one scene as fp32 accumulators in shared memory; per tile it stages g[c, tile] (coalesced 128-bit loads), and the
// thread sitting on the FIRST entry of each run of equal targets sums the run from the staged tile and adds it to the
// accumulator -- a plain read-modify-write, because inside a tile every target belongs to exactly one thread and tiles
// are separated by __syncthreads.  Rows are written out once, coalesced.  No atomics touch global memory; HBM traffic is
// the algorithmic minimum plus 6 bytes of sort output per position per CHUNK of channels.
constexpr int kBwdTile = 2048;
constexpr int kBwdSortThreads = 512;
constexpr int kBwdThreads = 512;

// grid (tiles_per_scene, b); dynamic smem: (n + 1) ints of bins + 32 ints of scan scratch
__global__ void __launch_bounds__(kBwdSortThreads) group_bwd_sort_kernel(const int *__restrict__ idx, int per, int n,
                                                                        int *__restrict__ skey, unsigned short *__restrict__ tperm) {
  extern __shared__ int s_bins[];  // [n] counts -> running cursors, then [32] warp partials
  int *wsum = s_bins + n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t base = (size_t)blockIdx.y * per + (size_t)blockIdx.x * kBwdTile;
  const int tc = min(kBwdTile, per - blockIdx.x * kBwdTile);
  for (int i = tid; i < n; i += kBwdSortThreads) s_bins[i] = 0;
  __syncthreads();
  int key[kBwdTile / kBwdSortThreads];
#pragma unroll
  for (int r = 0; r < kBwdTile / kBwdSortThreads; ++r) {
    const int e = r * kBwdSortThreads + tid;
    int k = -1;
    if (e < tc) {
      k = __ldg(idx + base + e);
      if ((unsigned)k >= (unsigned)n) k = -1;  // out-of-range indices are dropped (undefined behaviour in the reference)
      else atomicAdd(&s_bins[k], 1);
    }
    key[r] = k;
  }
  __syncthreads();
  // exclusive scan of s_bins[0..n): each thread owns a contiguous chunk
  const int chunk = (n + kBwdSortThreads - 1) / kBwdSortThreads;
  const int c0 = min(n, tid * chunk), c1 = min(n, c0 + chunk);
  int local = 0;
  for (int i = c0; i < c1; ++i) local += s_bins[i];
  int incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kBwdSortThreads / 32 ? wsum[lane] : 0;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += o;
    }
    wsum[lane] = wi - w;  // exclusive prefix of the warp totals; lane 31 slot unused beyond 16 warps
    if (lane == kBwdSortThreads / 32 - 1) wsum[31] = wi;  // total number of valid entries
  }
  __syncthreads();
  int run = wsum[warp] + incl - local;
  const int total = wsum[31];
  for (int i = c0; i < c1; ++i) {
    const int cnt = s_bins[i];
    s_bins[i] = run;
    run += cnt;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kBwdTile / kBwdSortThreads; ++r) {
    const int e = r * kBwdSortThreads + tid;
    if (key[r] >= 0) {
      const int pos = atomicAdd(&s_bins[key[r]], 1);
      skey[base + pos] = key[r];
      tperm[base + pos] = (unsigned short)e;
    }
  }
  for (int e = total + tid; e < tc; e += kBwdSortThreads) skey[base + e] = -1, tperm[base + e] = 0;
}

// Shared-memory stage of the accumulate kernel: one tile of CC gradient rows + its sort output.
template <int CC>
struct BwdStage {
  float gt[CC][kBwdTile];
  int sk[kBwdTile];
  unsigned short tp[kBwdTile];
};

// phase A + B of one tile (see the comment above): `acc` [CC][n] accumulators, stage `st`, tc valid entries
template <int CC>
__device__ __forceinline__ void bwd_process_tile(float *acc, int n, const BwdStage<CC> &st, int tc, int *contk, float *contv) {
  constexpr int kSteps = kBwdTile / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // phase A: one sorted entry per lane, segmented (by target) warp reduction, run heads update the accumulators
  for (int stp = warp; stp < kSteps; stp += kBwdThreads / 32) {
    const int o = stp * 32 + lane;
    const int k = o < tc ? st.sk[o] : -1;
    const int p = st.tp[o];
    float v[CC];
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) v[cc] = k >= 0 ? st.gt[cc][p] : 0.f;
    const int prev = __shfl_up_sync(0xffffffffu, k, 1);
    const bool head = (lane == 0) || (prev != k);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned after = lane == 31 ? 0u : (heads >> (lane + 1));
    const int run = after ? (__ffs(after) - 1) : (31 - lane);
    if (heads != 0xffffffffu) {  // some run is longer than one entry
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
          const float other = __shfl_down_sync(0xffffffffu, v[cc], d);
          if (d <= run) v[cc] += other;
        }
      }
    }
    const bool cont = lane == 0 && o > 0 && k >= 0 && st.sk[o - 1] == k;  // my run started in the previous warp step
    if (lane == 0) {
      contk[stp] = cont ? k : -1;
      if (cont) {
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) contv[cc * kSteps + stp] = v[cc];
      }
    }
    if (head && k >= 0 && !cont) {
#pragma unroll
      for (int cc = 0; cc < CC; ++cc) acc[(size_t)cc * n + k] += v[cc];
    }
  }
  __syncthreads();
  // phase B: fold the continuation partials (ascending targets), passes of 32 by warp 0
  if (warp == 0) {
#pragma unroll
    for (int pass = 0; pass < kSteps / 32; ++pass) {
      const int e = pass * 32 + lane;
      const int k = contk[e];
      if (__any_sync(0xffffffffu, k >= 0)) {
        float v[CC];
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) v[cc] = k >= 0 ? contv[cc * kSteps + e] : 0.f;
        const int key = k >= 0 ? k : -2 - lane;  // unique keys for empty slots so they never merge
        const int prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (prev != key);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const unsigned after = lane == 31 ? 0u : (heads >> (lane + 1));
        const int run = after ? (__ffs(after) - 1) : (31 - lane);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
            const float other = __shfl_down_sync(0xffffffffu, v[cc], d);
            if (d <= run) v[cc] += other;
          }
        }
        if (head && k >= 0) {
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) acc[(size_t)cc * n + k] += v[cc];
        }
      }
      __syncwarp();
    }
  }
}

// grid b * chunks.  dynamic smem: BwdStage<CC>[2] (double buffered, filled by the TMA engine) | contk[64] | contv[CC][64] |
// acc[CC][n].  bulk_ok: per % 8 == 0 and 16-byte aligned pointers (cp.async.bulk); otherwise tiles are loaded by the threads.
template <int CC>
__global__ void __launch_bounds__(kBwdThreads) group_bwd_accum_kernel(const float *__restrict__ grad_out, const int *__restrict__ skey,
                                                                     const unsigned short *__restrict__ tperm,
                                                                     float *__restrict__ grad_points, int c, int n, int per,
                                                                     int chunks, int bulk_ok) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  constexpr int kSteps = kBwdTile / 32;
  BwdStage<CC> *stage = reinterpret_cast<BwdStage<CC> *>(s_raw);
  int *contk = reinterpret_cast<int *>(stage + 2);
  float *contv = reinterpret_cast<float *>(contk + kSteps);
  float *acc = contv + CC * kSteps;
  __shared__ uint64_t full[2];
  const int tid = threadIdx.x;
  const int scene = blockIdx.x / chunks, chunk = blockIdx.x - scene * chunks;
  const int ch0 = chunk * CC;
  const int nch = min(CC, c - ch0);
  for (int i = tid; i < CC * n; i += kBwdThreads) acc[i] = 0.f;
  for (int i = tid; i < 2 * (int)(sizeof(BwdStage<CC>) / 4); i += kBwdThreads) reinterpret_cast<int *>(stage)[i] = 0;  // rows >= nch stay 0
  const float *g = grad_out + ((size_t)scene * c + ch0) * per;
  const size_t sbase = (size_t)scene * per;
  const int ntiles = (per + kBwdTile - 1) / kBwdTile;

  if (bulk_ok) {
    if (tid == 0) {
      mbar_init(&full[0], 1);
      mbar_init(&full[1], 1);
      fence_mbar_init();
    }
    fence_proxy_async();  // the zero fill above (generic proxy) is ordered before the bulk writes (async proxy)
    __syncthreads();
    auto issue = [&](int t) {
      const int e0 = t * kBwdTile, tc = min(kBwdTile, per - e0);
      BwdStage<CC> &st = stage[t & 1];
      mbar_arrive_expect_tx(&full[t & 1], (uint32_t)tc * (4u * nch + 6u));
      for (int cc = 0; cc < nch; ++cc) bulk_g2s(st.gt[cc], g + (size_t)cc * per + e0, (uint32_t)tc * 4u, &full[t & 1]);
      bulk_g2s(st.sk, skey + sbase + e0, (uint32_t)tc * 4u, &full[t & 1]);
      bulk_g2s(st.tp, tperm + sbase + e0, (uint32_t)tc * 2u, &full[t & 1]);
    };
    if (tid == 0) issue(0);
    for (int t = 0; t < ntiles; ++t) {
      if (tid == 0 && t + 1 < ntiles) issue(t + 1);  // buffer (t+1)&1 was released by the barrier ending tile t-1
      mbar_wait(&full[t & 1], (uint32_t)((t >> 1) & 1));
      bwd_process_tile<CC>(acc, n, stage[t & 1], min(kBwdTile, per - t * kBwdTile), contk, contv);
      __syncthreads();
    }
  } else {
    for (int t = 0; t < ntiles; ++t) {
      const int e0 = t * kBwdTile, tc = min(kBwdTile, per - e0);
      BwdStage<CC> &st = stage[0];
      __syncthreads();
      for (int cc = 0; cc < nch; ++cc)
        for (int i = tid; i < tc; i += kBwdThreads) st.gt[cc][i] = __ldg(g + (size_t)cc * per + e0 + i);
      for (int i = tid; i < tc; i += kBwdThreads) st.sk[i] = __ldg(skey + sbase + e0 + i), st.tp[i] = __ldg(tperm + sbase + e0 + i);
      __syncthreads();
      bwd_process_tile<CC>(acc, n, st, tc, contk, contv);
    }
  }
  __syncthreads();
  for (int cc = 0; cc < nch; ++cc) {
    float *dst = grad_points + ((size_t)scene * c + ch0 + cc) * n;
    for (int i = tid; i < n; i += kBwdThreads) dst[i] += acc[(size_t)cc * n + i];  // "+=": the entry point accumulates
  }
}

__global__ void group_bwd_generic_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, float *__restrict__ grad_points,
                                         int c, int n, size_t per, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / per, pos = e - row * per;
    const size_t scene = row / c;
    atomicAdd(grad_points + row * n + __ldg(idx + scene * per + pos), __ldg(grad_out + e));
  }
}

// ---- gather (C x m, tiny) ------------------------------------------------------------------------------------------
__global__ void gather_fwd_kernel(const float *__restrict__ points, const int *__restrict__ idx, float *__restrict__ out, int c, int n,
                                  int m, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / m, j = e - row * m;
    const size_t scene = row / c;
    out[e] = __ldg(points + row * n + __ldg(idx + scene * m + j));
  }
}
__global__ void gather_bwd_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, float *__restrict__ grad_points, int c,
                                  int n, int m, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / m, j = e - row * m;
    const size_t scene = row / c;
    atomicAdd(grad_points + row * n + __ldg(idx + scene * m + j), __ldg(grad_out + e));
  }
}

static inline unsigned grid_for(size_t total, int threads) {
  size_t g = (total + threads - 1) / threads;
  const size_t cap = (size_t)num_sms() * 32;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <int V>
static int launch_group_fwd(const float *points, const int *idx, float *out, int b, int c, int n, size_t per, int CH, cudaStream_t s) {
  auto kern = group_fwd_kernel<V>;
  const size_t smem = (size_t)CH * n * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int chunks = (c + CH - 1) / CH;
  const int per4 = (int)(per / 4);
  const long long total = (long long)b * chunks * per4;
  int ctas_per_sm = (int)((220u * 1024u) / (smem + 1024));
  ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm);
  long long ctas = (long long)num_sms() * ctas_per_sm;
  if (g_tuning.group_split > 0) ctas *= g_tuning.group_split;
  const long long min_w = 2 * kGroupThreads;  // do not cut finer than two sweeps of the block
  if (ctas * min_w > total) ctas = (total + min_w - 1) / min_w;
  if (ctas < 1) ctas = 1;
  const long long wpc = (total + ctas - 1) / ctas;
  ctas = (total + wpc - 1) / wpc;
  kern<<<(unsigned)ctas, kGroupThreads, smem, s>>>(points, idx, out, c, n, per4, CH, chunks, total, wpc, (g_tuning.group_mode & 1) ? 0 : 1);
  count_launch();
  return finish_launch();
}

}  // namespace gb

using namespace gb;

extern "C" int gb_group_fwd(const float *points, const int *idx, float *out, int b, int c, int n, int npoints, int nsample,
                            gb_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || npoints < 0 || nsample < 0 || !points || !idx || !out) return (int)cudaErrorInvalidValue;
  const size_t per = (size_t)npoints * nsample;
  if (b == 0 || c == 0 || per == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const bool aligned = (per % 4 == 0) && (((uintptr_t)idx | (uintptr_t)out) & 15u) == 0 && per / 4 < (1u << 30);
  const size_t row_bytes = (size_t)n * sizeof(float);
  const size_t big = 200u * 1024u;
  if (aligned && row_bytes <= big && !(g_tuning.group_mode & 2)) {
    // channels per fill: as many as fit (V-interleaved), aiming at <= ~100 KB so that two CTAs share an SM when rows are small
    int V = c >= 4 && 4 * row_bytes <= big ? 4 : (c >= 2 && 2 * row_bytes <= big ? 2 : 1);
    const size_t budget = (size_t)V * row_bytes > 100u * 1024u ? big : 100u * 1024u;
    int CH = (int)(budget / row_bytes);
    CH -= CH % V;
    if (CH > ((c + V - 1) / V) * V) CH = ((c + V - 1) / V) * V;
    if (CH > 64) CH = 64;
    if (V == 4) return launch_group_fwd<4>(points, idx, out, b, c, n, per, CH, s);
    if (V == 2) return launch_group_fwd<2>(points, idx, out, b, c, n, per, CH, s);
    return launch_group_fwd<1>(points, idx, out, b, c, n, per, CH, s);
  }
  const size_t total = (size_t)b * c * per;
  group_fwd_generic_kernel<<<grid_for(total, 256), 256, 0, s>>>(points, idx, out, c, n, per, total);
  count_launch();
  return finish_launch();
}

// stream-ordered scratch from the device's default pool; the pool is told once to keep freed memory instead of
// returning it to the driver at every synchronisation (the default), which would make every call pay a fresh allocation
static cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    configured = true;
  }
  return cudaMallocAsync(p, bytes, s);
}

template <int CC>
static int launch_group_bwd_sorted(const float *grad_out, const int *skey, const unsigned short *tperm, float *grad_points, int b, int c,
                                   int n, size_t per, size_t smem, int vec_ok, cudaStream_t s) {
  auto kern = group_bwd_accum_kernel<CC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int chunks = (c + CC - 1) / CC;
  kern<<<(unsigned)(b * chunks), kBwdThreads, smem, s>>>(grad_out, skey, tperm, grad_points, c, n, (int)per, chunks, vec_ok);
  count_launch();
  return finish_launch();
}

static size_t bwd_accum_smem(int CC, int n) {
  const size_t stage = (size_t)CC * kBwdTile * sizeof(float) + kBwdTile * (sizeof(int) + sizeof(unsigned short));
  return 2 * stage + (kBwdTile / 32) * (sizeof(int) + CC * sizeof(float)) + (size_t)CC * n * sizeof(float);
}

extern "C" int gb_group_bwd(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int npoints,
                            int nsample, gb_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || npoints < 0 || nsample < 0 || !grad_out || !idx || !grad_points) return (int)cudaErrorInvalidValue;
  const size_t per = (size_t)npoints * nsample;
  if (b == 0 || c == 0 || per == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;

  // ---- sorted, atomic-free path: worth the one-off sort when there are enough channels to amortise it ----
  const size_t sort_smem = ((size_t)n + 32) * sizeof(int);
  if (!(g_tuning.group_mode & 4) && c >= 4 && sort_smem <= 200u * 1024u && bwd_accum_smem(1, n) <= 200u * 1024u && b <= 65535 &&
      per < (1u << 30) && (size_t)b * per < (1u << 31)) {
    // channels per CTA: as many as keep two CTAs on an SM, but enough CTAs to fill the chip
    int CC = 8;
    while (CC > 1 && (bwd_accum_smem(CC, n) > 110u * 1024u || (long)b * ((c + CC - 1) / CC) < (long)num_sms())) CC >>= 1;
    const size_t tiles = (per + kBwdTile - 1) / kBwdTile;
    int *skey = nullptr;
    const size_t key_bytes = ((size_t)b * per * sizeof(int) + 255) & ~(size_t)255;
    cudaError_t e = scratch_alloc((void **)&skey, key_bytes + (size_t)b * per * sizeof(unsigned short), s);
    if (e != cudaSuccess) return (int)e;
    unsigned short *tperm = reinterpret_cast<unsigned short *>(reinterpret_cast<unsigned char *>(skey) + key_bytes);
    e = cudaFuncSetAttribute(group_bwd_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem);
    int rc = (int)e;
    if (!rc) {
      group_bwd_sort_kernel<<<dim3((unsigned)tiles, b), kBwdSortThreads, sort_smem, s>>>(idx, (int)per, n, skey, tperm);
      count_launch();
      rc = finish_launch();
    }
    const int vec_ok = (per % 8 == 0) && (((uintptr_t)grad_out & 15u) == 0);  // bulk (TMA) tile loads possible
    if (!rc) {
      const size_t smem = bwd_accum_smem(CC, n);
      switch (CC) {
        case 8: rc = launch_group_bwd_sorted<8>(grad_out, skey, tperm, grad_points, b, c, n, per, smem, vec_ok, s); break;
        case 4: rc = launch_group_bwd_sorted<4>(grad_out, skey, tperm, grad_points, b, c, n, per, smem, vec_ok, s); break;
        case 2: rc = launch_group_bwd_sorted<2>(grad_out, skey, tperm, grad_points, b, c, n, per, smem, vec_ok, s); break;
        default: rc = launch_group_bwd_sorted<1>(grad_out, skey, tperm, grad_points, b, c, n, per, smem, vec_ok, s); break;
      }
    }
    cudaFreeAsync(skey, s);
    return rc;
  }

  const bool aligned = (per % 4 == 0) && (((uintptr_t)idx | (uintptr_t)grad_out) & 15u) == 0 && (size_t)b * c <= 65535 && per / 4 < (1u << 30);
  if (aligned) {
    const int per4 = (int)(per / 4);
    dim3 grid((per4 + 1023) / 1024, b * c);
    group_bwd_kernel<<<grid, 256, 0, s>>>(grad_out, idx, grad_points, c, n, per4);
  } else {
    const size_t total = (size_t)b * c * per;
    group_bwd_generic_kernel<<<grid_for(total, 256), 256, 0, s>>>(grad_out, idx, grad_points, c, n, per, total);
  }
  count_launch();
  return finish_launch();
}

extern "C" int gb_gather_fwd(const float *points, const int *idx, float *out, int b, int c, int n, int m, gb_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || m < 0 || !points || !idx || !out) return (int)cudaErrorInvalidValue;
  const size_t total = (size_t)b * c * m;
  if (total == 0) return 0;
  gather_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(points, idx, out, c, n, m, total);
  count_launch();
  return finish_launch();
}

extern "C" int gb_gather_bwd(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int m,
                             gb_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || m < 0 || !grad_out || !idx || !grad_points) return (int)cudaErrorInvalidValue;
  const size_t total = (size_t)b * c * m;
  if (total == 0) return 0;
  gather_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(grad_out, idx, grad_points, c, n, m, total);
  count_launch();
  return finish_launch();
}
