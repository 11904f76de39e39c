// scatter.cu -- atomic-free segmented scatter-add, the backward of grouping_operation and of three_interpolate.
//
// Replaces group_points_grad_kernel (PointNet/_ext_src/src/group_points_gpu.cu:69-90), group_points_grad_kernel_fast
// (pointnet2_batch/src/group_points_gpu.cu:9-22), three_interpolate_grad_kernel (interpolate_gpu.cu:121-148) and
// three_interpolate_grad_kernel_fast (pointnet2_batch/src/interpolate_gpu.cu:127-149): one float atomicAdd per element
// (group) or three (interpolate) into global memory -- the L2 atomic units retire about one lane per SM per cycle, which
// is what bounds the reference.
//
// Both ops are   grad[b,c,key[b,e]] += src[b,c,e / DIV] * (w[b,e])   with DIV = 1 (group) or 3 (interpolate): the
// targets depend on e only, not on the channel.  So:
//   pass 1 (once per call, shared by all channels): every tile of T consecutive entries is counting-sorted by target in
//     shared memory; the output is one packed word per entry, (tile-local entry << 16) | target, in target order
//     (+ the weights in the same order for the interpolate backward);
//   pass 2: a CTA owns CC channels of one scene with a [n][CC] fp32 accumulator in shared memory.  Tiles of the CC source
//     rows stream in through the TMA engine (cp.async.bulk + mbarrier, up to 4 stages); each warp owns 64 consecutive
//     sorted entries of the tile, one per lane per step: runs of equal targets are summed with a segmented warp
//     reduction whose depth adapts to the longest run in the step (none when all 32 targets differ), and the head of
//     each run does a PLAIN vector read-modify-write of the accumulator -- inside a tile a target belongs to one warp,
//     except for the run that crosses a warp's boundary, whose two owners use shared-memory atomics (<= 2 targets per
//     warp per tile).  Tiles are separated by one __syncthreads.  Rows are written once, coalesced.
// HBM traffic is the algorithmic minimum (src read once, grad written once); the sort output (4-8 B per entry) is re-read
// from L2 once per channel chunk.
#include "common.cuh"

namespace gb {

constexpr int kSegTileStride = 2048;  // words of sort output per tile (tiles are padded to this)
constexpr int kSegSortThreads = 512;
constexpr int kSegThreads = 1024;  // 32 warps x 64 entries = one tile
constexpr int kSegMaxStages = 4;
constexpr unsigned kSegInvalid = 0xFFFFFFFFu;  // target 0xFFFF: dropped

// entries per tile: DIV = 3 keeps tiles on point boundaries and the TMA row pieces 16-byte multiples (680 * 4 B)
__host__ __device__ constexpr int seg_tile(int div) { return div == 3 ? 2040 : 2048; }

// grid (tiles, b); dynamic smem (n + 32) ints + 2048 u16.  idx [b, per] -> packed [b, tiles, 2048] (+ wsorted, same
// layout, one spare tile per scene) and bnd [b, tiles + 1, 64]
template <bool WEIGHTED>
__global__ void __launch_bounds__(kSegSortThreads) seg_sort_kernel(const int *__restrict__ idx, const float *__restrict__ weight,
                                                                   int per, int n, int T, unsigned *__restrict__ packed,
                                                                   float *__restrict__ wsorted, unsigned char *__restrict__ bnd) {
  extern __shared__ int s_bins[];  // [n] counts -> running cursors, then [32] warp partials, then [2048] sorted keys (u16)
  int *wsum = s_bins + n;
  unsigned short *skeys = reinterpret_cast<unsigned short *>(wsum + 32);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t ibase = (size_t)blockIdx.y * per + (size_t)blockIdx.x * T;
  const size_t obase = ((size_t)blockIdx.y * (gridDim.x + 1) + blockIdx.x) * kSegTileStride;  // one spare tile per scene
  const int tc = min(T, per - (int)blockIdx.x * T);
  for (int i = tid; i < n; i += kSegSortThreads) s_bins[i] = 0;
  __syncthreads();
  constexpr int R = kSegTileStride / kSegSortThreads;
  int key[R];
  float wv[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * kSegSortThreads + tid;
    int k = -1;
    wv[r] = 0.f;
    if (e < tc) {
      k = __ldg(idx + ibase + e);
      if (WEIGHTED) wv[r] = __ldg(weight + ibase + e);
      if ((unsigned)k >= (unsigned)n) k = -1;  // out-of-range targets are dropped (undefined behaviour in the reference)
      else atomicAdd(&s_bins[k], 1);
    }
    key[r] = k;
  }
  __syncthreads();
  // exclusive scan of s_bins[0..n): each thread owns a contiguous chunk
  const int chunk = (n + kSegSortThreads - 1) / kSegSortThreads;
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
    const int w = lane < kSegSortThreads / 32 ? wsum[lane] : 0;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += o;
    }
    wsum[lane] = wi - w;  // exclusive prefix of the warp totals (16 warps)
    if (lane == kSegSortThreads / 32 - 1) wsum[31] = wi;  // number of valid entries
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
  for (int r = 0; r < R; ++r) {
    const int e = r * kSegSortThreads + tid;
    if (key[r] >= 0) {
      const int pos = atomicAdd(&s_bins[key[r]], 1);
      packed[obase + pos] = ((unsigned)e << 16) | (unsigned)key[r];
      skeys[pos] = (unsigned short)key[r];
      if (WEIGHTED) wsorted[obase + pos] = wv[r];
    }
  }
  for (int e = total + tid; e < kSegTileStride; e += kSegSortThreads) {
    packed[obase + e] = kSegInvalid;
    skeys[e] = 0xFFFFu;
    if (WEIGHTED) wsorted[obase + e] = 0.f;
  }
  __syncthreads();
  // boundary bytes for the accumulate kernel: boundary w (between its warps w-1 and w) moves from entry 64 w to the next
  // run head within 32 entries (low 6 bits = shift); bit 7 = there is none, the run stays split (shared target)
  if (tid <= 32) {  // boundaries 0 and 32 are the tile's ends: always 0
    unsigned char code = 0;
    const int p = tid * 64;
    if (tid > 0 && tid < 32 && skeys[p] != 0xFFFFu && skeys[p] == skeys[p - 1]) {
      int sh = 1;
      while (sh <= 32 && skeys[p + sh] == skeys[p]) ++sh;  // p + 32 < 2048
      code = sh <= 32 ? (unsigned char)sh : (unsigned char)0x80u;
    }
    bnd[((size_t)blockIdx.y * (gridDim.x + 1) + blockIdx.x) * 64 + tid] = code;
  }
}

template <int CC>
__device__ __forceinline__ void seg_rmw(float *a, const float (&v)[CC]) {
  if (CC % 4 == 0) {
    float4 x[CC / 4 > 0 ? CC / 4 : 1];
#pragma unroll
    for (int q = 0; q < CC / 4; ++q) x[q] = reinterpret_cast<float4 *>(a)[q];
#pragma unroll
    for (int q = 0; q < CC / 4; ++q) {
      x[q].x += v[4 * q], x[q].y += v[4 * q + 1], x[q].z += v[4 * q + 2], x[q].w += v[4 * q + 3];
      reinterpret_cast<float4 *>(a)[q] = x[q];
    }
  } else if (CC == 2) {
    float2 x = *reinterpret_cast<float2 *>(a);
    x.x += v[0], x.y += v[1];
    *reinterpret_cast<float2 *>(a) = x;
  } else {
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) a[cc] += v[cc];
  }
}

// One step = 32 sorted entries at fixed positions, one per lane.  `mine`: the entry belongs to this warp (see the boundary
// bytes below); entries of other warps count as invalid.
//   * all 32 targets differ (the common case for large n): every lane does a plain vector read-modify-write;
//   * short runs of equal targets (rank of a lane inside its run <= 3): rank-by-rank passes, each conflict free;
//   * longer runs (a padded neighbourhood: up to nsample copies of one index): segmented warp reduction whose depth
//     adapts to the longest run, then the run heads update.
// Only when a run longer than 32 entries straddles this warp's boundary (`shared` targets) do its two owners fall back
// to shared-memory atomics for that one target.
template <int CC, int DIV, bool WEIGHTED>
__device__ __forceinline__ void seg_step(float *acc, const float *gt, int TP, unsigned pk, float w, bool mine, int lane, unsigned lemask,
                                         bool any_shared, unsigned kshare0, unsigned kshare1) {
  const unsigned k = mine ? (pk & 0xFFFFu) : 0xFFFFu;
  const unsigned tp = pk >> 16;
  const bool valid = k != 0xFFFFu;
  const unsigned sp = DIV == 3 ? (tp * 43691u) >> 17 : tp;  // tp / 3, exact for tp < 2^16
  float v[CC];
#pragma unroll
  for (int cc = 0; cc < CC; ++cc) {
    v[cc] = valid ? gt[cc * TP + sp] : 0.f;
    if (WEIGHTED) v[cc] = __fmul_rn(v[cc], w);  // the reference adds g * w_t (interpolate_gpu.cu:144-146)
  }
  float *a = acc + (size_t)k * CC;
  const unsigned prev = __shfl_up_sync(0xffffffffu, k, 1);
  const bool head = (lane == 0) || (prev != k);
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  if (any_shared) {  // rare: reduce fully, heads of shared targets use atomics
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
    if (head && valid) {
      if (k == kshare0 || k == kshare1) {
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) atomicAdd(a + cc, v[cc]);
      } else {
        seg_rmw<CC>(a, v);
      }
    }
    return;
  }
  if (heads == 0xffffffffu) {
    if (valid) seg_rmw<CC>(a, v);
    return;
  }
  const int rank = lane - (31 - __clz(heads & lemask));  // position of this lane inside its run
  const int maxrank = __reduce_max_sync(0xffffffffu, valid ? rank : 0);  // runs of masked / padding entries do not count
  if (maxrank <= 3) {
    for (int p = 0; p <= maxrank; ++p) {
      if (rank == p && valid) seg_rmw<CC>(a, v);
      __syncwarp();
    }
    return;
  }
  const unsigned after = lane == 31 ? 0u : (heads >> (lane + 1));
  const int run = after ? (__ffs(after) - 1) : (31 - lane);  // followers of this lane inside its run
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    if (d > maxrank) break;
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
      const float other = __shfl_down_sync(0xffffffffu, v[cc], d);
      if (d <= run) v[cc] += other;
    }
  }
  if (head && valid) seg_rmw<CC>(a, v);
}

// grid b * chunks, kSegThreads threads.  src [b,c,per_src] (per_src = entries / DIV); packed/wsorted [b,tiles+1,2048]
// (+32 words); bnd [b,tiles+1,64] boundary bytes; grad [b,c,n] (+= or =).  dynamic smem: stage[stages][CC][TP] | acc[n][CC].
//
// Warp w of the CTA owns the sorted entries [64 w + shift_w, 64 (w+1) + shift_{w+1}) of a tile: the sort kernel moved
// every 64-entry boundary right to the next run head (shift <= 32), so no run of equal targets is split between two
// warps and the accumulator needs no atomics.  Lanes sit on FIXED positions (three steps of 32 from 64 w; the third only
// when shift_{w+1} > 0) and mask the entries they do not own.
template <int CC, int DIV, bool WEIGHTED>
__global__ void __launch_bounds__(kSegThreads, 1) seg_accum_kernel(const float *__restrict__ src, const unsigned *__restrict__ packed,
                                                                   const float *__restrict__ wsorted,
                                                                   const unsigned char *__restrict__ bnd, float *__restrict__ grad,
                                                                   int c, int n, int per_src, int tiles, int chunks, int stages,
                                                                   int bulk_ok, int overwrite, size_t src_stride) {
  constexpr int T = seg_tile(DIV), TP = T / DIV;
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ uint64_t full[kSegMaxStages];
  float *stage = reinterpret_cast<float *>(s_raw);
  float *acc = stage + (size_t)stages * CC * TP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lemask = (2u << lane) - 1u;
  const int scene = blockIdx.x / chunks, chunk = blockIdx.x - scene * chunks;
  const int ch0 = chunk * CC;
  const int nch = min(CC, c - ch0);
  for (int i = tid; i < CC * n; i += kSegThreads) acc[i] = 0.f;
  if (nch < CC)  // rows of missing channels stay zero
    for (int i = tid; i < stages * CC * TP; i += kSegThreads) stage[i] = 0.f;
  const float *g = src + (size_t)scene * src_stride + (size_t)ch0 * per_src;
  // sort output of this scene; one spare tile behind the last one keeps the prefetch below unconditional
  const unsigned *pp = packed + (size_t)scene * (tiles + 1) * kSegTileStride + warp * 64 + lane;
  const float *wp = WEIGHTED ? wsorted + (size_t)scene * (tiles + 1) * kSegTileStride + warp * 64 + lane : nullptr;
  const unsigned char *bp = bnd + (size_t)scene * (tiles + 1) * 64 + warp;

  auto issue = [&](int t, int sidx) {  // thread 0: bulk copies of tile t's CC row pieces into stage sidx
    const int p0 = t * TP, pc = min(TP, per_src - p0);
    float *dst = stage + (size_t)sidx * CC * TP;
    mbar_arrive_expect_tx(&full[sidx], (uint32_t)pc * 4u * (uint32_t)nch);
    for (int cc = 0; cc < nch; ++cc) bulk_g2s(dst + cc * TP, g + (size_t)cc * per_src + p0, (uint32_t)pc * 4u, &full[sidx]);
  };
  if (bulk_ok) {
    if (tid == 0) {
      for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
      fence_mbar_init();
    }
    fence_proxy_async();  // the zero fill above (generic proxy) is ordered before the bulk writes (async proxy)
    __syncthreads();
    if (tid == 0)
      for (int t = 0; t < stages - 1 && t < tiles; ++t) issue(t, t);
  } else {
    __syncthreads();
  }

  // sort output of tile 0 (registers); the next tile's is fetched from L2 while the current one is processed
  unsigned pk0 = pp[0], pk1 = pp[32], pk2 = pp[64];
  float w0 = 0.f, w1 = 0.f, w2 = 0.f;
  if (WEIGHTED) w0 = wp[0], w1 = wp[32], w2 = wp[64];
  unsigned b0 = bp[0], b1 = bp[1];
  int sidx = 0, issue_sidx = stages - 1;  // stage of tile t / of tile t + stages - 1
  uint32_t parity = 0;
  for (int t = 0; t < tiles; ++t) {
    const unsigned c0 = pk0, c1 = pk1, c2 = pk2, cb0 = b0, cb1 = b1;
    const float cw0 = w0, cw1 = w1, cw2 = w2;
    pp += kSegTileStride, bp += 64;
    pk0 = pp[0], pk1 = pp[32], pk2 = pp[64];
    if (WEIGHTED) wp += kSegTileStride, w0 = wp[0], w1 = wp[32], w2 = wp[64];
    b0 = bp[0], b1 = bp[1];
    const float *gt = stage + (size_t)sidx * CC * TP;
    if (bulk_ok) {
      // the stage of tile t + stages - 1 was consumed by tile t - 1, which every thread has left (barrier below)
      if (tid == 0 && t + stages - 1 < tiles) issue(t + stages - 1, issue_sidx);
      mbar_wait(&full[sidx], parity);
    } else {
      const int p0 = t * TP, pc = min(TP, per_src - p0);
      float *dst = stage + (size_t)sidx * CC * TP;
      for (int cc = 0; cc < nch; ++cc)
        for (int i = tid; i < pc; i += kSegThreads) dst[cc * TP + i] = __ldg(g + (size_t)cc * per_src + p0 + i);
      __syncthreads();
    }
    // ownership: [64 w + shift0, 64 (w+1) + shift1)
    const int shift0 = cb0 & 63u, shift1 = cb1 & 63u;
    const bool any_shared = ((cb0 | cb1) & 0x80u) != 0u;
    unsigned kshare0 = 0xFFFFu, kshare1 = 0xFFFFu;
    if (any_shared) {  // a run of more than 32 entries straddles a boundary of this warp: its target is shared
      const unsigned kf = __shfl_sync(0xffffffffu, c0 & 0xFFFFu, 0), kl = __shfl_sync(0xffffffffu, c1 & 0xFFFFu, 31);
      if (cb0 & 0x80u) kshare0 = kf;
      if (cb1 & 0x80u) kshare1 = kl;
    }
    seg_step<CC, DIV, WEIGHTED>(acc, gt, TP, c0, cw0, lane >= shift0, lane, lemask, any_shared, kshare0, kshare1);
    __syncwarp();  // a run continuing into the next step updates the same accumulator
    seg_step<CC, DIV, WEIGHTED>(acc, gt, TP, c1, cw1, true, lane, lemask, any_shared, kshare0, kshare1);
    if (shift1) {
      __syncwarp();
      seg_step<CC, DIV, WEIGHTED>(acc, gt, TP, c2, cw2, lane < shift1, lane, lemask, any_shared, kshare0, kshare1);
    }
    __syncthreads();  // targets change owner between tiles; also releases the stage
    if (++sidx == stages) sidx = 0, parity ^= 1u;
    if (++issue_sidx == stages) issue_sidx = 0;
  }
  for (int cc = 0; cc < nch; ++cc) {
    float *dst = grad + ((size_t)scene * c + ch0 + cc) * n;
    if (overwrite) {
      for (int i = tid; i < n; i += kSegThreads) dst[i] = acc[(size_t)i * CC + cc];
    } else {  // the reference entry points accumulate into the caller's (zero-filled) tensor
      int i = tid;
      for (; i + 3 * kSegThreads < n; i += 4 * kSegThreads) {
        const float d0 = dst[i], d1 = dst[i + kSegThreads], d2 = dst[i + 2 * kSegThreads], d3 = dst[i + 3 * kSegThreads];
        dst[i] = d0 + acc[(size_t)i * CC + cc];
        dst[i + kSegThreads] = d1 + acc[(size_t)(i + kSegThreads) * CC + cc];
        dst[i + 2 * kSegThreads] = d2 + acc[(size_t)(i + 2 * kSegThreads) * CC + cc];
        dst[i + 3 * kSegThreads] = d3 + acc[(size_t)(i + 3 * kSegThreads) * CC + cc];
      }
      for (; i < n; i += kSegThreads) dst[i] += acc[(size_t)i * CC + cc];
    }
  }
}

// ---- dense mode: few targets (n <= 4096) -------------------------------------------------------------------------------
// When a tile of ~2048 entries hits every target about once or more (InvResMLP groupings with n = m <= 2048, the
// interpolation backward with <= 1024 known points), entry-parallel processing spends its time merging runs.  Here the
// roles flip: a THREAD owns a target (TPT targets when n > 1024) and CT channels for the whole kernel and keeps their sums
// in REGISTERS; per tile it walks its targets' slices of the sorted entry list and adds the staged source values.  No
// accumulator in shared memory, no atomics, no run merging; the only barrier releases the stage.  Which thread owns
// which target is decided per call from the targets' (sampled) degrees: see seg_perm_kernel.
//
// Sort output per tile ("blob", one bulk copy): tp[2048] u16 (tile-local entry of each sorted position) | start[NS] u16
// (first sorted position of every target, start[n] = number of valid entries; NS = n + 1 rounded up to 8).
// Entries per tile: 2048 (group) or 3 x 2040 (interpolate: 2040 points, 8160-byte row pieces for the bulk copies).
// MUL (1, 2 or 4) widens the tile when a thread holds fewer channels, keeping ~64 KB of source rows per stage: the more
// entries of a target a tile holds, the better the lanes of a warp are used (slice lengths are Poisson-like).
__host__ __device__ constexpr int seg_dense_tile(int div, int mul) { return mul * (div == 3 ? 6120 : 2048); }
__host__ __device__ constexpr int seg_dense_stride(int div, int mul) { return mul * (div == 3 ? 6144 : 2048); }  // tp slots per blob
__host__ __device__ inline int seg_dense_ns(int n) { return (n + 1 + 7) & ~7; }
__host__ __device__ inline size_t seg_dense_blob(int n, int div, int mul) {
  return (size_t)seg_dense_stride(div, mul) * 2 + (size_t)seg_dense_ns(n) * 2;
}

// grid (tiles, b); dynamic smem (n + 32) ints
// deg (optional) [b, n]: every deg_stride-th tile adds its per-target counts -- a sampled degree of each target, which
// seg_perm_kernel turns into the thread -> target assignment of the accumulate kernel.
template <int DIV, int MUL>
__global__ void __launch_bounds__(kSegSortThreads) seg_sort_dense_kernel(const int *__restrict__ idx, int per, int n,
                                                                         unsigned char *__restrict__ blobs, int *__restrict__ deg,
                                                                         int deg_stride) {
  constexpr int T = seg_dense_tile(DIV, MUL), kStride = seg_dense_stride(DIV, MUL);
  extern __shared__ int s_bins[];
  int *wsum = s_bins + n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t ibase = (size_t)blockIdx.y * per + (size_t)blockIdx.x * T;
  unsigned short *tp = reinterpret_cast<unsigned short *>(blobs + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * seg_dense_blob(n, DIV, MUL));
  unsigned short *start = tp + kStride;
  const int tc = min(T, per - (int)blockIdx.x * T);
  for (int i = tid; i < n; i += kSegSortThreads) s_bins[i] = 0;
  __syncthreads();
  constexpr int R = kStride / kSegSortThreads;
  int key[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * kSegSortThreads + tid;
    int k = -1;
    if (e < tc) {
      k = __ldg(idx + ibase + e);
      if ((unsigned)k >= (unsigned)n) k = -1;  // out-of-range targets are dropped
      else atomicAdd(&s_bins[k], 1);
    }
    key[r] = k;
  }
  __syncthreads();
  if (deg && blockIdx.x % deg_stride == 0) {
    int *dg = deg + (size_t)blockIdx.y * n;
    for (int i = tid; i < n; i += kSegSortThreads) {
      const int cnt = s_bins[i];
      if (cnt) atomicAdd(dg + i, cnt);
    }
  }
  const int chunk = (n + kSegSortThreads - 1) / kSegSortThreads;
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
    const int w = lane < kSegSortThreads / 32 ? wsum[lane] : 0;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += o;
    }
    wsum[lane] = wi - w;
    if (lane == kSegSortThreads / 32 - 1) wsum[31] = wi;
  }
  __syncthreads();
  int run = wsum[warp] + incl - local;
  const int total = wsum[31];
  for (int i = c0; i < c1; ++i) {
    const int cnt = s_bins[i];
    s_bins[i] = run;
    start[i] = (unsigned short)run;
    run += cnt;
  }
  for (int i = n + tid; i < seg_dense_ns(n); i += kSegSortThreads) start[i] = (unsigned short)total;
  __syncthreads();
  // Scatter into shared memory first (the cursor order is whatever the atomics made it), then every entry finds its RANK
  // among its target's entries by counting the smaller entry numbers of the slice: each slice leaves in ascending entry
  // order, so the sums of seg_dense_kernel are bit-reproducible run to run.  The scan costs the slice length per entry
  // (~6 for the interpolation backward, ~30 for a heavy ball-query target); slices beyond kStableMax entries of one tile
  // (degenerate inputs: every neighbourhood empty) keep the cursor order.
  constexpr int kStableMax = 256;
  unsigned short *s_tp = reinterpret_cast<unsigned short *>(wsum + 32);
  int pos[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    pos[r] = -1;
    if (key[r] >= 0) {
      pos[r] = atomicAdd(&s_bins[key[r]], 1);
      s_tp[pos[r]] = (unsigned short)(r * kSegSortThreads + tid);
    }
  }
  __syncthreads();  // s_bins[k] is now the END of target k's slice, i.e. the start of target k + 1's
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (key[r] < 0) continue;
    const int e = r * kSegSortThreads + tid;
    const int lo = key[r] ? s_bins[key[r] - 1] : 0, hi = s_bins[key[r]];
    int at = pos[r];
    if (hi - lo <= kStableMax) {
      at = lo;
      for (int i = lo; i < hi; ++i) at += (int)s_tp[i] < e;
    }
    tp[at] = (unsigned short)e;
  }
  for (int e = total + tid; e < kStride; e += kSegSortThreads) tp[e] = 0;
}

// Thread -> target assignment of the accumulate kernel.  A warp walks the slices of its 32 targets in lock step, so a
// tile costs it as many iterations as its LONGEST slice; the number of entries per target is far from uniform (ball-query
// neighbourhoods padded with their first hit, dense and sparse regions: degrees of 0..400 around a mean of 64), so with
// targets in index order most lanes idle (measured: 12 of 32 lanes active per instruction).  Sorting the targets by
// degree puts targets of similar degree in the same warp.  grid b, 1024 threads: counting sort of the n <= 4096 targets
// by (sampled) degree, descending; perm [b, n] u16 = target of slot s.  Ownership only: sums per target are unchanged.
constexpr int kSegDegBins = 1024;
__global__ void __launch_bounds__(kSegThreads) seg_perm_kernel(const int *__restrict__ deg, int n, unsigned short *__restrict__ perm) {
  __shared__ int s_hist[kSegDegBins];
  __shared__ int s_w[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  deg += (size_t)blockIdx.x * n;
  perm += (size_t)blockIdx.x * n;
  s_hist[tid] = 0;
  __syncthreads();
  int bin[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int t = r * kSegThreads + tid;
    bin[r] = -1;
    if (t < n) {
      bin[r] = kSegDegBins - 1 - min(deg[t], kSegDegBins - 1);  // bin 0 = highest degree
      atomicAdd(&s_hist[bin[r]], 1);
    }
  }
  __syncthreads();
  const int mine = s_hist[tid];
  int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = s_w[lane];
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += o;
    }
    s_w[lane] = wi - w;
  }
  __syncthreads();
  s_hist[tid] = s_w[warp] + incl - mine;  // exclusive prefix = first slot of the bin, then its cursor
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (bin[r] >= 0) perm[atomicAdd(&s_hist[bin[r]], 1)] = (unsigned short)(r * kSegThreads + tid);
}

// grid b * chunks, kSegThreads threads.  Thread -> target slot ts = tid % NT and channel group cg = tid / NT
// (NT = 2^nt_log2 >= n when n <= 1024, else 1024 with TPT targets ts, ts + 1024, ...); CC = CT * (1024 / NT) channels per
// CTA.  dynamic smem per stage: gt[CC][TP] f32 | wt[T] f32 (weighted) | blob.
template <int CT, int TPT, int DIV, bool WEIGHTED, int MUL>
__global__ void __launch_bounds__(kSegThreads, 1) seg_dense_kernel(const float *__restrict__ src, const unsigned char *__restrict__ blobs,
                                                                   const float *__restrict__ weight, float *__restrict__ grad, int c,
                                                                   int n, int per_src, int tiles, int chunks, int nt_log2, int stages,
                                                                   int stage_bytes, int bulk_ok, int overwrite, size_t src_stride,
                                                                   const unsigned short *__restrict__ perm) {
  constexpr int T = seg_dense_tile(DIV, MUL), TP = T / DIV;
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ uint64_t full[kSegMaxStages];
  const int tid = threadIdx.x;
  const int NT = 1 << nt_log2, G = kSegThreads >> nt_log2, CC = CT * G;
  const int ts = tid & (NT - 1), cg = tid >> nt_log2;
  const int scene = blockIdx.x / chunks, chunk = blockIdx.x - scene * chunks;
  const int ch0 = chunk * CC;
  const int nch = min(CC, c - ch0);
  const float *g = src + (size_t)scene * src_stride + (size_t)ch0 * per_src;
  const size_t blob = seg_dense_blob(n, DIV, MUL);
  const unsigned char *bsrc = blobs + (size_t)scene * tiles * blob;
  const float *wsrc = WEIGHTED ? weight + (size_t)scene * per_src * DIV : nullptr;
  const int wt_off = CC * TP * 4, blob_off = wt_off + (WEIGHTED ? T * 4 : 0);

  if (nch < CC)  // rows of missing channels stay zero
    for (int i = tid; i < stages * stage_bytes / 4; i += kSegThreads) reinterpret_cast<float *>(s_raw)[i] = 0.f;
  auto issue = [&](int t, int sidx) {
    const int p0 = t * TP, pc = min(TP, per_src - p0);
    unsigned char *dst = s_raw + (size_t)sidx * stage_bytes;
    mbar_arrive_expect_tx(&full[sidx], (uint32_t)pc * 4u * (uint32_t)nch + (WEIGHTED ? (uint32_t)pc * DIV * 4u : 0u) + (uint32_t)blob);
    for (int cc = 0; cc < nch; ++cc)
      bulk_g2s(dst + (size_t)cc * TP * 4, g + (size_t)cc * per_src + p0, (uint32_t)pc * 4u, &full[sidx]);
    if (WEIGHTED) bulk_g2s(dst + wt_off, wsrc + (size_t)p0 * DIV, (uint32_t)pc * DIV * 4u, &full[sidx]);
    bulk_g2s(dst + blob_off, bsrc + (size_t)t * blob, (uint32_t)blob, &full[sidx]);
  };
  if (bulk_ok) {
    if (tid == 0) {
      for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
      fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0)
      for (int t = 0; t < stages - 1 && t < tiles; ++t) issue(t, t);
  } else {
    __syncthreads();
  }

  float a[TPT][CT];
  int tgt[TPT];  // the targets this thread owns: slot ts + j * NT of the (degree-sorted) assignment; -1 = none
#pragma unroll
  for (int j = 0; j < TPT; ++j) {
    // degree-sorted assignment: serpentine over the TPT rows of NT slots, so a thread (and hence a warp) pairs a
    // high-degree target with a low-degree one and the warps of a CTA carry similar totals
    const int slot = perm && (j & 1) ? (j + 1) * NT - 1 - ts : ts + j * NT;
    tgt[j] = slot < n ? (perm ? (int)perm[(size_t)scene * n + slot] : slot) : -1;
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) a[j][cc] = 0.f;
  }

  int sidx = 0, issue_sidx = stages - 1;
  uint32_t parity = 0;
  for (int t = 0; t < tiles; ++t) {
    unsigned char *st = s_raw + (size_t)sidx * stage_bytes;
    if (bulk_ok) {
      if (tid == 0 && t + stages - 1 < tiles) issue(t + stages - 1, issue_sidx);
      mbar_wait(&full[sidx], parity);
    } else {
      const int p0 = t * TP, pc = min(TP, per_src - p0);
      for (int cc = 0; cc < nch; ++cc)
        for (int i = tid; i < pc; i += kSegThreads) reinterpret_cast<float *>(st)[cc * TP + i] = __ldg(g + (size_t)cc * per_src + p0 + i);
      if (WEIGHTED)
        for (int i = tid; i < pc * DIV; i += kSegThreads) reinterpret_cast<float *>(st + wt_off)[i] = __ldg(wsrc + (size_t)p0 * DIV + i);
      for (int i = tid; i < (int)(blob / 4); i += kSegThreads)
        reinterpret_cast<unsigned *>(st + blob_off)[i] = __ldg(reinterpret_cast<const unsigned *>(bsrc + (size_t)t * blob) + i);
      __syncthreads();
    }
    const float *gt = reinterpret_cast<const float *>(st) + (size_t)cg * CT * TP;
    const float *wt = reinterpret_cast<const float *>(st + wt_off);
    const unsigned short *tp = reinterpret_cast<const unsigned short *>(st + blob_off);
    const unsigned short *start = tp + seg_dense_stride(DIV, MUL);
#pragma unroll
    for (int j = 0; j < TPT; ++j) {
      const int k = tgt[j];
      if (k >= 0) {
        const int e1 = start[k + 1];
        for (int i = start[k]; i < e1; ++i) {
          const unsigned p = tp[i];
          const unsigned sp = DIV == 3 ? (p * 43691u) >> 17 : p;
          const float wv = WEIGHTED ? wt[p] : 1.f;
#pragma unroll
          for (int cc = 0; cc < CT; ++cc) {
            const float v = gt[cc * TP + sp];
            a[j][cc] += WEIGHTED ? __fmul_rn(v, wv) : v;  // the reference adds g * w_t (interpolate_gpu.cu:144-146)
          }
        }
      }
    }
    __syncthreads();  // releases the stage
    if (++sidx == stages) sidx = 0, parity ^= 1u;
    if (++issue_sidx == stages) issue_sidx = 0;
  }
#pragma unroll
  for (int j = 0; j < TPT; ++j) {
    const int k = tgt[j];
    if (k >= 0) {
#pragma unroll
      for (int cc = 0; cc < CT; ++cc) {
        const int ch = cg * CT + cc;
        if (ch < nch) {
          float *dst = grad + ((size_t)scene * c + ch0 + ch) * n + k;
          *dst = overwrite ? a[j][cc] : *dst + a[j][cc];
        }
      }
    }
  }
}

// stream-ordered scratch from the device's default pool; the pool is told once to keep freed memory instead of
// returning it to the driver at every synchronisation (the default), which would make every call pay a fresh allocation
cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t s) {
  static std::atomic<bool> configured[kMaxDevices];  // per device: one process may drive every GPU of the box
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevices && !configured[dev].load(std::memory_order_acquire)) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);  // idempotent: a race sets it twice
    }
    configured[dev].store(true, std::memory_order_release);
  }
  return cudaMallocAsync(p, bytes, s);
}

constexpr size_t kSegSmemBudget = 227u * 1024u - 1024u;  // dynamic; the mbarriers are static

template <int CC, int DIV, bool WEIGHTED>
static int launch_seg_accum(const float *src, size_t src_stride, const unsigned *packed, const float *wsorted, const unsigned char *bnd,
                            float *grad, int b, int c, int n, int per_src, int tiles, int bulk_ok, int overwrite, cudaStream_t s) {
  constexpr int TP = seg_tile(DIV) / DIV;
  const size_t acc_bytes = (size_t)CC * n * sizeof(float), stage_bytes = (size_t)CC * TP * sizeof(float);
  int stages = (int)((kSegSmemBudget - acc_bytes) / stage_bytes);
  stages = stages > kSegMaxStages ? kSegMaxStages : stages;
  if (stages > tiles) stages = tiles < 1 ? 1 : tiles;
  if (!bulk_ok) stages = 1;
  const size_t smem = acc_bytes + stages * stage_bytes;
  auto kern = seg_accum_kernel<CC, DIV, WEIGHTED>;
  if (int rc_ = raise_smem_limit(kern, smem)) return rc_;
  const int chunks = (c + CC - 1) / CC;
  kern<<<(unsigned)(b * chunks), kSegThreads, smem, s>>>(src, packed, wsorted, bnd, grad, c, n, per_src, tiles, chunks, stages, bulk_ok,
                                                         overwrite, src_stride);
  count_launch();
  return finish_launch();
}

template <int CT, int TPT, int DIV, bool WEIGHTED, int MUL>
static int launch_seg_dense(const float *src, size_t src_stride, const unsigned char *blobs, const float *weight, float *grad, int b, int c,
                            int n, int per_src, int tiles, int nt_log2, int bulk_ok, int overwrite, const unsigned short *perm,
                            cudaStream_t s) {
  constexpr int T = seg_dense_tile(DIV, MUL), TP = T / DIV;
  const int G = kSegThreads >> nt_log2, CC = CT * G;
  const size_t stage_bytes = ((size_t)CC * TP * 4 + (WEIGHTED ? (size_t)T * 4 : 0) + seg_dense_blob(n, DIV, MUL) + 127) & ~(size_t)127;
  int stages = (int)(kSegSmemBudget / stage_bytes);
  stages = stages > kSegMaxStages ? kSegMaxStages : stages;
  if (stages > tiles) stages = tiles;
  if (!bulk_ok || stages < 1) stages = 1;
  const size_t smem = stages * stage_bytes;
  auto kern = seg_dense_kernel<CT, TPT, DIV, WEIGHTED, MUL>;
  if (int rc_ = raise_smem_limit(kern, smem)) return rc_;
  const int chunks = (c + CC - 1) / CC;
  kern<<<(unsigned)(b * chunks), kSegThreads, smem, s>>>(src, blobs, weight, grad, c, n, per_src, tiles, chunks, nt_log2, stages,
                                                         (int)stage_bytes, bulk_ok, overwrite, src_stride, perm);
  count_launch();
  return finish_launch();
}

constexpr int kSegDenseMaxN = 4096;

// dense mode (thread-owned targets, register accumulators) for n <= 4096
static int seg_scatter_dense(const float *src, size_t src_stride, const int *key, const float *weight, float *grad, int b, int c, int n,
                             size_t entries, int div, int overwrite, cudaStream_t s) {
  const int per_src = (int)(entries / div);
  int nt_log2 = 8;  // >= 256 target slots: at most 4 channel groups per CTA, so one channel per thread always fits
  while ((1 << nt_log2) < n && nt_log2 < 10) ++nt_log2;
  if (g_tuning.scatter_nt >= 8 && g_tuning.scatter_nt <= 10 && (n + (1 << g_tuning.scatter_nt) - 1) >> g_tuning.scatter_nt <= 4)
    nt_log2 = g_tuning.scatter_nt;
  const int NT = 1 << nt_log2, G = kSegThreads / NT;
  const int tpt = (n + NT - 1) / NT;  // 1 when n <= 1024, else 2..4
  // tile multiplier of a channel count: ~64 KB of source rows per stage (group); the interpolate tile is 2040 points
  auto mul_of = [&](int ct) { return div == 3 ? 1 : (ct * G >= 8 ? 1 : (ct * G >= 4 ? 2 : 4)); };
  auto stage_of = [&](int ct) {
    const int T = seg_dense_tile(div, mul_of(ct)), TP = T / div;
    return (size_t)ct * G * TP * 4 + (weight ? (size_t)T * 4 : 0) + seg_dense_blob(n, div, mul_of(ct)) + 128;
  };
  // channels per thread: as many as registers (CT * TPT <= 16), two stages of shared memory and the grid allow; with two
  // or more targets per thread, 4 channels x a 4096-entry tile beat 8 x 2048 (B200, n = 2048: 830 vs 893 us)
  int CT = tpt >= 2 ? 4 : 8;
  while (CT > 1 && (CT * (tpt > 2 ? 4 : tpt) > 16 || CT * G > ((c + 3) & ~3) * 2 || 2 * stage_of(CT) > kSegSmemBudget ||
                    (long)b * ((c + CT * G - 1) / (CT * G)) < (long)num_sms()))
    CT >>= 1;
  if (g_tuning.scatter_cc == 1 || g_tuning.scatter_cc == 2 || g_tuning.scatter_cc == 4 || g_tuning.scatter_cc == 8) {
    CT = g_tuning.scatter_cc;
    while (CT > 1 && (CT * (tpt > 2 ? 4 : tpt) > 16 || 2 * stage_of(CT) > kSegSmemBudget)) CT >>= 1;
  }
  const int mul = mul_of(CT);
  const int T = seg_dense_tile(div, mul);
  const int tiles = (int)((entries + T - 1) / T);
  const size_t blob = seg_dense_blob(n, div, mul);
  // degree-sorted thread -> target assignment (see seg_perm_kernel) when a scene has enough tiles for the imbalance to
  // matter; degrees are sampled from every deg_stride-th tile (at least 8 tiles per scene)
  const bool balance = tiles >= 8 && !(g_tuning.scatter_mode & 4);
  const int deg_stride = tiles >= 32 ? tiles / 8 : (tiles >= 16 ? 2 : 1);
  const size_t blobs_bytes = ((size_t)b * tiles * blob + 255) & ~(size_t)255;
  const size_t deg_bytes = balance ? (((size_t)b * n * sizeof(int) + 255) & ~(size_t)255) : 0;
  const size_t perm_bytes = balance ? (size_t)b * n * sizeof(unsigned short) : 0;
  unsigned char *blobs = nullptr;
  cudaError_t e = scratch_alloc((void **)&blobs, blobs_bytes + deg_bytes + perm_bytes, s);
  if (e != cudaSuccess) return (int)e;
  int *deg = balance ? reinterpret_cast<int *>(blobs + blobs_bytes) : nullptr;
  unsigned short *perm = balance ? reinterpret_cast<unsigned short *>(blobs + blobs_bytes + deg_bytes) : nullptr;
  if (balance) {
    e = cudaMemsetAsync(deg, 0, (size_t)b * n * sizeof(int), s);
    if (e != cudaSuccess) {
      cudaFreeAsync(blobs, s);
      return (int)e;
    }
  }
  const size_t sort_smem = ((size_t)n + 64) * sizeof(int) + (size_t)seg_dense_stride(div, mul) * sizeof(unsigned short);
  const dim3 sgrid((unsigned)tiles, b);
  int rc = 0;
#define GB_SORT_DENSE(DIVV, MULV)                                                                                        \
  do {                                                                                                                   \
    rc = raise_smem_limit(seg_sort_dense_kernel<DIVV, MULV>, sort_smem);                                                 \
    if (!rc) seg_sort_dense_kernel<DIVV, MULV><<<sgrid, kSegSortThreads, sort_smem, s>>>(key, (int)entries, n, blobs, deg, deg_stride); \
  } while (0)
  if (div == 3) GB_SORT_DENSE(3, 1);
  else if (mul == 1) GB_SORT_DENSE(1, 1);
  else if (mul == 2) GB_SORT_DENSE(1, 2);
  else GB_SORT_DENSE(1, 4);
#undef GB_SORT_DENSE
  if (rc) {
    cudaFreeAsync(blobs, s);
    return rc;
  }
  count_launch();
  rc = finish_launch();
  if (!rc && balance) {
    seg_perm_kernel<<<b, kSegThreads, 0, s>>>(deg, n, perm);
    count_launch();
    rc = finish_launch();
  }
  if (!rc) {
    // bulk copies: 16-byte aligned row pieces (and weight pieces: per_src * div * 4 bytes per scene)
    const int bulk_ok = (per_src % 4 == 0) && (((uintptr_t)src & 15u) == 0) && (!weight || ((uintptr_t)weight & 15u) == 0) &&
                        2 * stage_of(CT) <= kSegSmemBudget;
#define GB_DENSE_ARGS src, src_stride, blobs, weight, grad, b, c, n, per_src, tiles, nt_log2, bulk_ok, overwrite, perm, s
#define GB_DENSE_CASE(CTV, TPTV)                                                                      \
  do {                                                                                                \
    if (div == 3) rc = launch_seg_dense<CTV, TPTV, 3, true, 1>(GB_DENSE_ARGS);                        \
    else if (mul == 1) rc = launch_seg_dense<CTV, TPTV, 1, false, 1>(GB_DENSE_ARGS);                  \
    else if (mul == 2) rc = launch_seg_dense<CTV, TPTV, 1, false, 2>(GB_DENSE_ARGS);                  \
    else rc = launch_seg_dense<CTV, TPTV, 1, false, 4>(GB_DENSE_ARGS);                                \
  } while (0)
    if (tpt <= 1) {
      switch (CT) {
        case 8: GB_DENSE_CASE(8, 1); break;
        case 4: GB_DENSE_CASE(4, 1); break;
        case 2: GB_DENSE_CASE(2, 1); break;
        default: GB_DENSE_CASE(1, 1); break;
      }
    } else if (tpt <= 2) {
      switch (CT) {
        case 8: GB_DENSE_CASE(8, 2); break;
        case 4: GB_DENSE_CASE(4, 2); break;
        case 2: GB_DENSE_CASE(2, 2); break;
        default: GB_DENSE_CASE(1, 2); break;
      }
    } else {
      switch (CT) {
        case 4: GB_DENSE_CASE(4, 4); break;
        case 2: GB_DENSE_CASE(2, 4); break;
        default: GB_DENSE_CASE(1, 4); break;
      }
    }
#undef GB_DENSE_CASE
#undef GB_DENSE_ARGS
  }
  cudaFreeAsync(blobs, s);
  return rc;
}

bool seg_scatter_supported(int b, int c, int n, size_t entries, int div) {
  const int TP = seg_tile(div) / div;
  return c >= 4 && n < 0xFFFF && ((size_t)n + 32) * sizeof(int) + 4096 <= 200u * 1024u &&
         (size_t)n * sizeof(float) + 2 * (size_t)TP * sizeof(float) <= kSegSmemBudget && b <= 65535 && entries < (1u << 30) &&
         (size_t)b * entries < (1u << 31);
}

// grad[b,c,key[b,e]] += src[b,c,e/div] * (weight ? weight[b,e] : 1); key/weight [b,entries]; src [b,c,entries/div].
// Two shapes are instantiated: div = 1 without weights (group) and div = 3 with weights (interpolate).
// overwrite != 0: grad is fully written (no zero fill needed) instead of accumulated into.
int seg_scatter_add(const float *src, size_t src_stride, const int *key, const float *weight, float *grad, int b, int c, int n,
                    size_t entries, int div, int overwrite, cudaStream_t s) {
  if ((div != 1 && div != 3) || (div == 3) != (weight != nullptr)) return (int)cudaErrorInvalidValue;
  if (n <= kSegDenseMaxN && !(g_tuning.scatter_mode & 1)) return seg_scatter_dense(src, src_stride, key, weight, grad, b, c, n, entries, div, overwrite, s);
  const int T = seg_tile(div), TP = T / div;
  const int per_src = (int)(entries / div);
  const int tiles = (int)((entries + T - 1) / T);
  // per scene one spare (never processed, only prefetched) tile, plus 32 words behind each array for the third-step read
  const size_t words = (size_t)b * (tiles + 1) * kSegTileStride + 32;
  unsigned *packed = nullptr;
  cudaError_t e = scratch_alloc((void **)&packed, words * sizeof(unsigned) * (weight ? 2 : 1) + (size_t)b * (tiles + 1) * 64, s);
  if (e != cudaSuccess) return (int)e;
  float *wsorted = weight ? reinterpret_cast<float *>(packed + words) : nullptr;
  unsigned char *bnd = reinterpret_cast<unsigned char *>(packed + words * (weight ? 2 : 1));
  const size_t sort_smem = ((size_t)n + 32) * sizeof(int) + kSegTileStride * sizeof(unsigned short);
  int rc;
  if (weight) {
    rc = raise_smem_limit(seg_sort_kernel<true>, sort_smem);
    if (!rc) seg_sort_kernel<true><<<dim3((unsigned)tiles, b), kSegSortThreads, sort_smem, s>>>(key, weight, (int)entries, n, T, packed, wsorted, bnd);
  } else {
    rc = raise_smem_limit(seg_sort_kernel<false>, sort_smem);
    if (!rc) seg_sort_kernel<false><<<dim3((unsigned)tiles, b), kSegSortThreads, sort_smem, s>>>(key, nullptr, (int)entries, n, T, packed, nullptr, bnd);
  }
  if (!rc) {
    count_launch();
    rc = finish_launch();
  }
  if (!rc) {
    // channels per CTA: as many as fit beside two stages, but keep at least one CTA per SM
    int CC = 8;
    while (CC > 1 && ((size_t)CC * n * sizeof(float) + 2 * (size_t)CC * TP * sizeof(float) > kSegSmemBudget ||
                      (long)b * ((c + CC - 1) / CC) < (long)num_sms()))
      CC >>= 1;
    if (g_tuning.scatter_cc == 1 || g_tuning.scatter_cc == 2 || g_tuning.scatter_cc == 4 || g_tuning.scatter_cc == 8) {
      CC = g_tuning.scatter_cc;
      while (CC > 1 && (size_t)CC * n * sizeof(float) + 2 * (size_t)CC * TP * sizeof(float) > kSegSmemBudget) CC >>= 1;
    }
    // bulk (TMA) row pieces need 16-byte aligned global addresses and sizes
    const int bulk_ok = (per_src % 4 == 0) && (((uintptr_t)src & 15u) == 0);
#define GB_SEG_CASE(CCV)                                                                                            \
  case CCV:                                                                                                         \
    rc = div == 3 ? launch_seg_accum<CCV, 3, true>(src, src_stride, packed, wsorted, bnd, grad, b, c, n, per_src, tiles, bulk_ok, overwrite, s)  \
                  : launch_seg_accum<CCV, 1, false>(src, src_stride, packed, wsorted, bnd, grad, b, c, n, per_src, tiles, bulk_ok, overwrite, s); \
    break;
    switch (CC) {
      GB_SEG_CASE(8)
      GB_SEG_CASE(4)
      GB_SEG_CASE(2)
      default:
        GB_SEG_CASE(1)
    }
#undef GB_SEG_CASE
  }
  cudaFreeAsync(packed, s);
  return rc;
}

}  // namespace gb
