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
//   * large launches: a CTA stages the rows of ONE (scene, chunk) pair and sweeps ~384-512 KB of its output (aligned
//     partition: one fill per CTA, many short waves that hide each other's fills); small launches: the (scene, chunk) x
//     position work space is flattened and cut into equal contiguous ranges, one per resident CTA, so all 148 SMs finish
//     together whatever the shape.
// Backward is a scatter-add.  With >= 4 channels it runs as the atomic-free sorted segmented sum of scatter.cu; the
// kernels kept here serve few-channel calls (the xyz rows): coalesced 128-bit reads of grad_out and idx,
// red.global.add.f32 into the (L2-resident) gradient rows, with WARP-AGGREGATION of runs of equal indices first --
// ball/cylinder query pads a neighbourhood with copies of its first hit, so sparse neighbourhoods collapse to one
// atomic per run.
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

// ---- grouped coordinates of QueryAndGroup / CylinderQueryAndGroup in one pass ------------------------------------------
// out[b, :, j, k] = ((xyz[b, idx[b,j,k]] - new_xyz[b,j]) * scale) . R[b,j]: what pointnet2_utils.py:178-190,281-291 build
// from transpose + grouping_operation + subtract + divide + permute + matmul + permute (seven passes over [B,3,m,ns] and a
// batched 3x3 sgemm).  Each op rounds as the torch op it replaces: fp32 subtract, multiply by the fp32 reciprocal of the
// radius (ATen divides by a CPU scalar that way), dot product accumulated over k = 0, 1, 2 with fma.
// VEC = 4: one thread per 4 consecutive samples of a query (nsample % 4 == 0), 128-bit idx load and stores.
// Items t0, t0 + tstep, ... < total of the flattened (scene, query, sample group) space.
template <bool ROT, int VEC>
__device__ __forceinline__ void group_xyz_items(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                                const int *__restrict__ idx, const float *__restrict__ rot, float *__restrict__ out,
                                                int n, int m, int ns, float scale, int use_scale, size_t out_stride, size_t total,
                                                size_t t0, size_t tstep) {
  const size_t per = (size_t)m * ns;
  const int nsv = ns / VEC;
  for (size_t t = t0; t < total; t += tstep) {
    const size_t qrow = t / nsv;  // scene * m + j
    const int kq = (int)(t - qrow * nsv);
    const size_t scene = qrow / m;
    const float qx = __ldg(new_xyz + qrow * 3), qy = __ldg(new_xyz + qrow * 3 + 1), qz = __ldg(new_xyz + qrow * 3 + 2);
    float r[9];
    if (ROT) {
#pragma unroll
      for (int e = 0; e < 9; ++e) r[e] = __ldg(rot + qrow * 9 + e);
    }
    int id[VEC];
    if (VEC == 4) {
      const int4 v = ld_nc_i4(idx + qrow * ns + (size_t)kq * 4);
      id[0] = v.x, id[VEC > 1 ? 1 : 0] = v.y, id[VEC > 2 ? 2 : 0] = v.z, id[VEC > 3 ? 3 : 0] = v.w;
    } else {
      id[0] = __ldg(idx + qrow * ns + kq);
    }
    float o[3][VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float *p = xyz + (scene * n + (size_t)id[e]) * 3;
      float dx = __fsub_rn(__ldg(p), qx), dy = __fsub_rn(__ldg(p + 1), qy), dz = __fsub_rn(__ldg(p + 2), qz);
      if (use_scale) dx = __fmul_rn(dx, scale), dy = __fmul_rn(dy, scale), dz = __fmul_rn(dz, scale);
      if (ROT) {
        o[0][e] = __fmaf_rn(dz, r[6], __fmaf_rn(dy, r[3], __fmul_rn(dx, r[0])));
        o[1][e] = __fmaf_rn(dz, r[7], __fmaf_rn(dy, r[4], __fmul_rn(dx, r[1])));
        o[2][e] = __fmaf_rn(dz, r[8], __fmaf_rn(dy, r[5], __fmul_rn(dx, r[2])));
      } else {
        o[0][e] = dx, o[1][e] = dy, o[2][e] = dz;
      }
    }
    float *dst = out + scene * out_stride + (qrow - scene * m) * ns + (size_t)kq * VEC;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      if (VEC == 4) st_cs_f4(dst + ch * per, make_float4(o[ch][0], o[ch][VEC > 1 ? 1 : 0], o[ch][VEC > 2 ? 2 : 0], o[ch][VEC > 3 ? 3 : 0]));
      else dst[ch * per] = o[ch][0];
    }
  }
}

// The grouped coordinates of a grouper folded into its feature-grouping launch (gb_group_xyz_feat): every CTA of
// group_fwd_kernel first takes its grid-stride share of the (scene, query, 4 samples) items -- ~3 / C of the launch's
// output, gathered straight from the L2-resident xyz -- so a QueryAndGroup is ONE launch instead of a 10-17 us
// latency-bound coordinate launch in front of the bandwidth-bound one.  total = 0: no fold.
struct XyzFold {
  const float *xyz, *new_xyz;
  float *out;
  int m, ns;
  float scale;
  int use_scale;
  size_t out_stride, total;
};

// points [b,c,n]; idx [b,per]; out [b,c,per]; per % 4 == 0.  CH = channels staged per fill (multiple of V),
// chunks = ceil(c / CH), per4 = per / 4, wpc = work (quads) per CTA.
template <int V>
__global__ void __launch_bounds__(kGroupThreads) group_fwd_kernel(const float *__restrict__ points, const int *__restrict__ idx,
                                                                 float *__restrict__ out, int c, int n, int per4, int CH, int chunks,
                                                                 long long total, long long wpc, int streaming, size_t out_stride,
                                                                 int ranges, const XyzFold fold) {
  extern __shared__ __align__(16) float s_rows[];  // [CH/V][n][V]
  using Vec = typename VecT<V>::type;
  Vec *srow = reinterpret_cast<Vec *>(s_rows);
  const int tid = threadIdx.x;
  if (fold.total)
    group_xyz_items<false, 4>(fold.xyz, fold.new_xyz, idx, nullptr, fold.out, n, fold.m, fold.ns, fold.scale, fold.use_scale,
                              fold.out_stride, fold.total, (size_t)blockIdx.x * kGroupThreads + tid, (size_t)gridDim.x * kGroupThreads);
  long long w = (long long)blockIdx.x * wpc;
  long long wend = min(total, w + wpc);
  if (ranges > 0) {  // aligned partition: CTA = (pair, r), the r-th of `ranges` equal position ranges of ONE (scene, chunk) pair
    const long long pair = blockIdx.x / ranges;
    const int r = (int)(blockIdx.x - pair * ranges);
    w = pair * per4 + ((long long)per4 * r) / ranges;
    wend = pair * per4 + ((long long)per4 * (r + 1)) / ranges;
  }
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
    int4 id_next = make_int4(0, 0, 0, 0);
    if (q0 + tid < q1) id_next = ld_nc_i4(ip + (size_t)(q0 + tid) * 4);
    for (int q = q0 + tid; q < q1; q += kGroupThreads) {
      // the next quad's indices (an L2 round trip) are fetched under this quad's gathers and stores (B200, n = m = 2048,
      // C = 128, ns = 64: 381 -> 366 us = 92 % of HBM peak; two quads ahead: no further gain)
      const int4 id = id_next;
      if (q + kGroupThreads < q1) id_next = ld_nc_i4(ip + (size_t)(q + kGroupThreads) * 4);
      for (int g = 0; g < gcount; ++g) {
        const Vec *row = srow + (size_t)g * n;
        const Vec a0 = row[id.x], a1 = row[id.y], a2 = row[id.z], a3 = row[id.w];
        const int ch0 = ch_base + g * V;
#pragma unroll
        for (int e = 0; e < V; ++e) {
          if (ch0 + e < c) {
            float *dst = out + (size_t)scene * out_stride + (size_t)(ch0 + e) * per + (size_t)q * 4;
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

// Large clouds (a source row of n floats no longer fits four or two at a time): ONE channel row per buffer, two buffers,
// filled by the TMA engine (one cp.async.bulk per row: the row is contiguous in global memory) while the previous row is
// swept -- fill and sweep overlap inside a CTA instead of across co-resident CTAs.  Work unit = (row, position segment);
// a CTA takes a contiguous run of units, so consecutive units usually share the staged row.
// points [b,c,n]; idx [b,per]; out rows `per` floats long, scenes out_stride apart.  dynamic smem: 2 * n floats.
__global__ void __launch_bounds__(kGroupThreads) group_fwd_rows_kernel(const float *__restrict__ points, const int *__restrict__ idx,
                                                                      float *__restrict__ out, int c, int n, int per4, int segs,
                                                                      int seg_len4, long long units, long long upc, size_t out_stride) {
  extern __shared__ __align__(128) float s_row[];  // [2][n]
  __shared__ uint64_t full[2];
  const int tid = threadIdx.x;
  const long long u0 = (long long)blockIdx.x * upc, u1 = min(units, u0 + upc);
  if (u0 >= u1) return;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const size_t per = (size_t)per4 * 4;
  const long long r_first = u0 / segs, r_last = (u1 - 1) / segs;
  auto load_row = [&](long long r, int buf) {  // thread 0
    mbar_arrive_expect_tx(&full[buf], (uint32_t)n * 4u);
    bulk_g2s(s_row + (size_t)buf * n, points + (size_t)r * n, (uint32_t)n * 4u, &full[buf]);
  };
  if (tid == 0) load_row(r_first, 0);
  long long cur_row = -1;
  int k = -1;  // index of the current row within this CTA: buffer k & 1, wait parity (k >> 1) & 1
  for (long long u = u0; u < u1; ++u) {
    const long long r = u / segs;
    const int sgm = (int)(u - r * segs);
    if (r != cur_row) {
      __syncthreads();  // every thread has left the previous row: its predecessor's buffer may be refilled
      cur_row = r, ++k;
      if (tid == 0 && r < r_last) load_row(r + 1, (k + 1) & 1);
      mbar_wait(&full[k & 1], (uint32_t)((k >> 1) & 1));
    }
    const float *row = s_row + (size_t)(k & 1) * n;
    const size_t scene = (size_t)(r / c), ch = (size_t)(r - (long long)scene * c);
    const int *ip = idx + scene * per;
    float *op = out + scene * out_stride + ch * per;
    const int q1 = min(per4, (sgm + 1) * seg_len4);
    int q = sgm * seg_len4 + tid;
    // four index loads (L2 latency) in flight per thread before the first gather
    for (; q + 3 * kGroupThreads < q1; q += 4 * kGroupThreads) {
      int4 id[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) id[u] = ld_nc_i4(ip + (size_t)(q + u * kGroupThreads) * 4);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        st_cs_f4(op + (size_t)(q + u * kGroupThreads) * 4, make_float4(row[id[u].x], row[id[u].y], row[id[u].z], row[id[u].w]));
    }
    for (; q < q1; q += kGroupThreads) {
      const int4 id = ld_nc_i4(ip + (size_t)q * 4);
      st_cs_f4(op + (size_t)q * 4, make_float4(row[id.x], row[id.y], row[id.z], row[id.w]));
    }
  }
}

// generic fallback (any shape / alignment): one thread per output element, gathers straight from global/L2
__global__ void group_fwd_generic_kernel(const float *__restrict__ points, const int *__restrict__ idx, float *__restrict__ out, int c,
                                         int n, size_t per, size_t total, size_t out_stride) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / per, pos = e - row * per;
    const size_t scene = row / c, ch = row - scene * c;
    out[scene * out_stride + ch * per + pos] = __ldg(points + row * n + __ldg(idx + scene * per + pos));
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
                                                        float *__restrict__ grad_points, int c, int n, int per4, size_t go_stride) {
  const size_t row = blockIdx.y;
  const size_t scene = row / c;
  const size_t per = (size_t)per4 * 4;
  const float *g = grad_out + scene * go_stride + (row - scene * c) * per;
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

// Few channels (c < 4: the grouped coordinates), where neither the warp-private nor the sorted path applies and the launch
// is latency-bound: ONE launch, one CTA per output row with the row's n sums in shared memory (shared-memory atomics, no
// zero fill of the output, no global atomics).  grid b*c, 1024 threads, dynamic smem n floats.
__global__ void __launch_bounds__(1024) group_bwd_row_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx,
                                                             float *__restrict__ grad_points, int c, int n, size_t per, size_t go_stride,
                                                             int overwrite, int vec) {
  extern __shared__ __align__(16) float s_acc[];
  const size_t row = blockIdx.x, scene = row / c;
  const float *g = grad_out + scene * go_stride + (row - scene * c) * per;
  const int *ip = idx + scene * per;
  for (int i = threadIdx.x; i < n; i += 1024) s_acc[i] = 0.f;
  __syncthreads();
  if (vec) {
    // four quads per thread in flight (the launch is latency-bound: loads first, then the updates); adjacent equal targets
    // (padded rows) are summed before the update
    constexpr int U = 4;
    const size_t quads = per / 4;
    for (size_t q0 = threadIdx.x; q0 < quads; q0 += (size_t)1024 * U) {
      int4 id[U];
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t q = q0 + (size_t)u * 1024;
        const bool ok = q < quads;
        id[u] = ok ? ld_nc_i4(ip + q * 4) : make_int4(-1, -1, -1, -1);
        v[u] = ok ? ld_nc_na_f4(g + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float acc = v[u].x;
        if (id[u].y == id[u].x) acc += v[u].y;
        else { if ((unsigned)id[u].x < (unsigned)n) atomicAdd(s_acc + id[u].x, acc); acc = v[u].y; }
        if (id[u].z == id[u].y) acc += v[u].z;
        else { if ((unsigned)id[u].y < (unsigned)n) atomicAdd(s_acc + id[u].y, acc); acc = v[u].z; }
        if (id[u].w == id[u].z) acc += v[u].w;
        else { if ((unsigned)id[u].z < (unsigned)n) atomicAdd(s_acc + id[u].z, acc); acc = v[u].w; }
        if ((unsigned)id[u].w < (unsigned)n) atomicAdd(s_acc + id[u].w, acc);
      }
    }
  } else {
    for (size_t e = threadIdx.x; e < per; e += 1024) {
      const int t = __ldg(ip + e);
      if ((unsigned)t < (unsigned)n) atomicAdd(s_acc + t, __ldg(g + e));
    }
  }
  __syncthreads();
  float *dst = grad_points + row * n;
  for (int i = threadIdx.x; i < n; i += 1024) dst[i] = overwrite ? s_acc[i] : dst[i] + s_acc[i];
}

__global__ void group_bwd_generic_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, float *__restrict__ grad_points,
                                         int c, int n, size_t per, size_t total, size_t go_stride) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / per, pos = e - row * per;
    const size_t scene = row / c;
    atomicAdd(grad_points + row * n + __ldg(idx + scene * per + pos), __ldg(grad_out + scene * go_stride + (row - scene * c) * per + pos));
  }
}

// ---- grouped coordinates as a launch of their own (no features, rotated crops, unaligned shapes): group_xyz_items above ----
template <bool ROT, int VEC>
__global__ void __launch_bounds__(256) group_xyz_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                                        const int *__restrict__ idx, const float *__restrict__ rot,
                                                        float *__restrict__ out, int n, int m, int ns, float scale, int use_scale,
                                                        size_t out_stride, size_t total) {
  group_xyz_items<ROT, VEC>(xyz, new_xyz, idx, rot, out, n, m, ns, scale, use_scale, out_stride, total,
                            blockIdx.x * (size_t)blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
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
static int launch_group_fwd(const float *points, const int *idx, float *out, int b, int c, int n, size_t per, int CH, size_t out_stride,
                            cudaStream_t s, const XyzFold &fold) {
  auto kern = group_fwd_kernel<V>;
  const size_t smem = (size_t)CH * n * sizeof(float);
  if (int rc_ = raise_smem_limit(kern, smem)) return rc_;
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
  long long wpc = (total + ctas - 1) / ctas;
  ctas = (total + wpc - 1) / wpc;
  // Aligned partition (default when a pair is large enough): every CTA stages the rows of ONE (scene, chunk) pair once and
  // sweeps 1/ranges of its positions, ~target_kb of output per CTA -- small enough that several waves of CTAs overlap
  // each other's fills and the tail is short, large enough that the fill (CH rows from L2) stays a small part.
  // (B200, 32 scenes: n = m = 2048, C = 128: 442 -> 379 us; n = 2048, m = 1024: 123 -> 111 us; n = m = 1024, C = 256: 226 -> 218 us;
  // the optimum moves by +-10 % with the CTA count, tests/ubench/fwd_shapes.py.)
  int ranges = 0;
  const long long pairs = (long long)b * chunks;
  if (!(g_tuning.group_mode & 16) && pairs * 1 <= (1LL << 30)) {
    const long long pair_out_kb = (long long)per4 * 16 * CH / 1024;
    const long long target_kb = g_tuning.group_target_kb > 0 ? (int)g_tuning.group_target_kb : (smem <= 64u * 1024u ? 512 : 384);
    long long r = (pair_out_kb + target_kb / 2) / target_kb;
    const long long max_r = per4 / min_w > 1 ? per4 / min_w : 1;
    r = r < 1 ? 1 : (r > max_r ? max_r : r);
    if (pairs * r >= 4 * ctas) {  // at least four waves (wave quantisation); smaller launches keep the flattened equal split
      ranges = (int)r;
      ctas = pairs * r;
      wpc = 0;
    }
  }
  kern<<<(unsigned)ctas, kGroupThreads, smem, s>>>(points, idx, out, c, n, per4, CH, chunks, total, wpc, (g_tuning.group_mode & 1) ? 0 : 1,
                                                   out_stride, ranges, fold);
  count_launch();
  return finish_launch();
}

}  // namespace gb

using namespace gb;

static int group_xyz_impl(const float *xyz, const float *new_xyz, const int *idx, const float *rot, float *out, int b, int n, int m,
                            int nsample, float scale, int use_scale, long long out_scene_stride, gb_stream_t stream) {
  if (b < 0 || n <= 0 || m < 0 || nsample < 0) return (int)cudaErrorInvalidValue;
  const size_t per = (size_t)m * nsample;
  if (b == 0 || per == 0) return 0;
  if (!xyz || !new_xyz || !idx || !out) return (int)cudaErrorInvalidValue;
  if (out_scene_stride < (long long)(3 * per)) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = (nsample % 4 == 0) && (out_scene_stride % 4 == 0) && ((((uintptr_t)idx | (uintptr_t)out) & 15u) == 0);
  const size_t total = vec ? (size_t)b * m * (nsample / 4) : (size_t)b * per;
  const unsigned grid = grid_for(total, 256);
  const size_t os = (size_t)out_scene_stride;
  if (vec) {
    if (rot) group_xyz_kernel<true, 4><<<grid, 256, 0, s>>>(xyz, new_xyz, idx, rot, out, n, m, nsample, scale, use_scale, os, total);
    else group_xyz_kernel<false, 4><<<grid, 256, 0, s>>>(xyz, new_xyz, idx, rot, out, n, m, nsample, scale, use_scale, os, total);
  } else {
    if (rot) group_xyz_kernel<true, 1><<<grid, 256, 0, s>>>(xyz, new_xyz, idx, rot, out, n, m, nsample, scale, use_scale, os, total);
    else group_xyz_kernel<false, 1><<<grid, 256, 0, s>>>(xyz, new_xyz, idx, rot, out, n, m, nsample, scale, use_scale, os, total);
  }
  count_launch();
  return finish_launch();
}

// fold: grouped coordinates to be produced by the same launch (staged path), or by a launch of their own in front of the
// other paths; nullptr = features only
static int group_fwd_impl(const float *points, const int *idx, float *out, int b, int c, int n, int npoints, int nsample,
                          long long out_scene_stride, gb_stream_t stream, const XyzFold *fold = nullptr) {
  if (b < 0 || c < 0 || n <= 0 || npoints < 0 || nsample < 0) return (int)cudaErrorInvalidValue;
  const size_t per = (size_t)npoints * nsample;
  if (b == 0 || c == 0 || per == 0) return 0;  // nothing to do (empty tensors have null data pointers)
  if (!points || !idx || !out) return (int)cudaErrorInvalidValue;
  if (out_scene_stride < (long long)((size_t)c * per)) return (int)cudaErrorInvalidValue;
  const size_t ostride = (size_t)out_scene_stride;
  cudaStream_t s = (cudaStream_t)stream;
  const bool aligned = (per % 4 == 0) && (ostride % 4 == 0) && (((uintptr_t)idx | (uintptr_t)out) & 15u) == 0 && per / 4 < (1u << 30);
  const size_t row_bytes = (size_t)n * sizeof(float);
  const size_t big = 200u * 1024u;
  XyzFold nofold;
  nofold.total = 0;
  const bool staged = aligned && row_bytes <= big && !(g_tuning.group_mode & 2) &&
                      !(4 * row_bytes > big && 2 * row_bytes <= big && n % 4 == 0 && (((uintptr_t)points & 15u) == 0) && !(g_tuning.group_mode & (2 | 8)));
  const bool foldable = fold && staged && !(g_tuning.group_mode & 32) && fold->ns % 4 == 0 && fold->out_stride % 4 == 0 &&
                        (((uintptr_t)fold->out) & 15u) == 0;
  if (fold && !foldable) {  // coordinates as a launch of their own
    if (int rc_ = group_xyz_impl(fold->xyz, fold->new_xyz, idx, nullptr, fold->out, b, n, fold->m, fold->ns, fold->scale, fold->use_scale,
                                 (long long)fold->out_stride, stream))
      return rc_;
  }
  const XyzFold &kf = foldable ? *fold : nofold;
  // rows too long to stage four channels at a time: TMA double-buffered single rows
  if (aligned && 4 * row_bytes > big && 2 * row_bytes <= big && n % 4 == 0 && (((uintptr_t)points & 15u) == 0) &&
      !(g_tuning.group_mode & (2 | 8))) {
    const int per4 = (int)(per / 4);
    const long long rows = (long long)b * c;
    const int grid = num_sms();
    int segs = (int)((4LL * grid + rows - 1) / rows);  // >= ~4 units per CTA so the contiguous split balances
    const int max_segs = per4 / (2 * kGroupThreads) > 1 ? per4 / (2 * kGroupThreads) : 1;
    segs = segs < 1 ? 1 : (segs > max_segs ? max_segs : segs);
    const int seg_len4 = (per4 + segs - 1) / segs;
    segs = (per4 + seg_len4 - 1) / seg_len4;
    const long long units = rows * segs;
    const long long ctas = units < grid ? units : grid;
    const long long upc = (units + ctas - 1) / ctas;
    const size_t smem = 2 * row_bytes;
    if (int rc_ = raise_smem_limit(group_fwd_rows_kernel, smem)) return rc_;
    group_fwd_rows_kernel<<<(unsigned)((units + upc - 1) / upc), kGroupThreads, smem, s>>>(points, idx, out, c, n, per4, segs, seg_len4, units,
                                                                                      upc, ostride);
    count_launch();
    return finish_launch();
  }
  if (aligned && row_bytes <= big && !(g_tuning.group_mode & 2)) {
    // channels per fill: as many as fit (V-interleaved), aiming at <= ~100 KB so that two CTAs share an SM when rows are small
    int V = c >= 4 && 4 * row_bytes <= big ? 4 : (c >= 2 && 2 * row_bytes <= big ? 2 : 1);
    const size_t budget = (size_t)V * row_bytes > 100u * 1024u ? big : 100u * 1024u;
    int CH = (int)(budget / row_bytes);
    CH -= CH % V;
    if (CH > ((c + V - 1) / V) * V) CH = ((c + V - 1) / V) * V;
    if (CH > 64) CH = 64;
    // short rows (n <= 1024): 16 channels per fill -- more co-resident CTAs and shorter fills beat fewer idx re-reads
    // (B200, 32 scenes, C = 256: n = 512: 72 -> 59 us, n = 1024: 218 -> 202 us with 512 KB ranges; tests/ubench/fwd_shapes.py)
    if (row_bytes <= 4096 && CH > 16 && V == 4) CH = 16;
    if (g_tuning.group_ch > 0 && g_tuning.group_ch % V == 0 && (size_t)g_tuning.group_ch * row_bytes <= big) CH = g_tuning.group_ch;
    if (V == 4) return launch_group_fwd<4>(points, idx, out, b, c, n, per, CH, ostride, s, kf);
    if (V == 2) return launch_group_fwd<2>(points, idx, out, b, c, n, per, CH, ostride, s, kf);
    return launch_group_fwd<1>(points, idx, out, b, c, n, per, CH, ostride, s, kf);
  }
  const size_t total = (size_t)b * c * per;
  group_fwd_generic_kernel<<<grid_for(total, 256), 256, 0, s>>>(points, idx, out, c, n, per, total, ostride);
  count_launch();
  return finish_launch();
}

extern "C" int gb_group_fwd(const float *points, const int *idx, float *out, int b, int c, int n, int npoints, int nsample,
                            gb_stream_t stream) {
  return group_fwd_impl(points, idx, out, b, c, n, npoints, nsample, (long long)c * npoints * nsample, stream);
}

extern "C" int gb_group_fwd_strided(const float *points, const int *idx, float *out, int b, int c, int n, int npoints, int nsample,
                                    long long out_scene_stride, gb_stream_t stream) {
  return group_fwd_impl(points, idx, out, b, c, n, npoints, nsample, out_scene_stride, stream);
}

static int group_bwd_impl(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int npoints, int nsample,
                          long long go_scene_stride, int overwrite, gb_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || npoints < 0 || nsample < 0) return (int)cudaErrorInvalidValue;
  const size_t per = (size_t)npoints * nsample;
  if (b > 0 && c > 0 && (!grad_points || (per > 0 && (!grad_out || !idx)))) return (int)cudaErrorInvalidValue;
  if (go_scene_stride < (long long)((size_t)c * per)) return (int)cudaErrorInvalidValue;
  const size_t gstride = (size_t)go_scene_stride;
  cudaStream_t s = (cudaStream_t)stream;
  if (b == 0 || c == 0) return 0;
  if (per == 0) return overwrite ? (int)cudaMemsetAsync(grad_points, 0, (size_t)b * c * n * sizeof(float), s) : 0;

  // ---- warp-private accumulators (scatter_private.cu): the backbone's shapes, no sort pass at all ----
  if (!(g_tuning.group_mode & 4) && scatter_private_supported(b, c, n, npoints, nsample, gstride, grad_out, idx))
    return scatter_private(grad_out, gstride, idx, grad_points, b, c, n, npoints, nsample, overwrite, s);
  // ---- sorted, atomic-free path (scatter.cu): worth the one-off sort when there are enough channels to amortise it ----
  if (!(g_tuning.group_mode & 4) && gstride % 4 == 0 && seg_scatter_supported(b, c, n, per, 1))
    return seg_scatter_add(grad_out, gstride, idx, nullptr, grad_points, b, c, n, per, 1, overwrite, s);
  const bool aligned16 = (per % 4 == 0) && (gstride % 4 == 0) && (((uintptr_t)idx | (uintptr_t)grad_out) & 15u) == 0;
  // ---- few channels, enough rows to occupy the GPU: one CTA per output row, sums in shared memory (B200, C = 3, 65536
  //      entries -> 20000 targets: 32 scenes 28 us against 56 us for memset + global atomics; 4 scenes 27 us against 17 us,
  //      hence the row count) ----
  if (c < 4 && (size_t)b * c >= 48 && (size_t)n * sizeof(float) <= 200u * 1024u && per <= (1u << 18) && !(g_tuning.group_mode & 8)) {
    const size_t smem = (size_t)n * sizeof(float);
    if (int rc_ = raise_smem_limit(group_bwd_row_kernel, smem)) return rc_;
    group_bwd_row_kernel<<<(unsigned)((size_t)b * c), 1024, smem, s>>>(grad_out, idx, grad_points, c, n, per, gstride, overwrite, aligned16 ? 1 : 0);
    count_launch();
    return finish_launch();
  }
  if (overwrite) {
    cudaError_t e = cudaMemsetAsync(grad_points, 0, (size_t)b * c * n * sizeof(float), s);
    if (e != cudaSuccess) return (int)e;
  }

  const bool aligned = aligned16 && (size_t)b * c <= 65535 && per / 4 < (1u << 30);
  if (aligned) {
    const int per4 = (int)(per / 4);
    dim3 grid((per4 + 1023) / 1024, b * c);
    group_bwd_kernel<<<grid, 256, 0, s>>>(grad_out, idx, grad_points, c, n, per4, gstride);
  } else {
    const size_t total = (size_t)b * c * per;
    group_bwd_generic_kernel<<<grid_for(total, 256), 256, 0, s>>>(grad_out, idx, grad_points, c, n, per, total, gstride);
  }
  count_launch();
  return finish_launch();
}

extern "C" int gb_group_bwd(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int npoints,
                            int nsample, gb_stream_t stream) {
  return group_bwd_impl(grad_out, idx, grad_points, b, c, n, npoints, nsample, (long long)c * npoints * nsample, 0, stream);
}

extern "C" int gb_group_bwd_set(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int npoints,
                                int nsample, gb_stream_t stream) {
  return group_bwd_impl(grad_out, idx, grad_points, b, c, n, npoints, nsample, (long long)c * npoints * nsample, 1, stream);
}

extern "C" int gb_group_bwd_strided(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int npoints,
                                    int nsample, long long grad_out_scene_stride, int overwrite, gb_stream_t stream) {
  return group_bwd_impl(grad_out, idx, grad_points, b, c, n, npoints, nsample, grad_out_scene_stride, overwrite ? 1 : 0, stream);
}

extern "C" int gb_gather_fwd(const float *points, const int *idx, float *out, int b, int c, int n, int m, gb_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || m < 0) return (int)cudaErrorInvalidValue;
  const size_t total = (size_t)b * c * m;
  if (total == 0) return 0;
  if (!points || !idx || !out) return (int)cudaErrorInvalidValue;
  gather_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(points, idx, out, c, n, m, total);
  count_launch();
  return finish_launch();
}

extern "C" int gb_gather_bwd(const float *grad_out, const int *idx, float *grad_points, int b, int c, int n, int m,
                             gb_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || m < 0) return (int)cudaErrorInvalidValue;
  const size_t total = (size_t)b * c * m;
  if (total == 0) return 0;
  if (!grad_out || !idx || !grad_points) return (int)cudaErrorInvalidValue;
  gather_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(grad_out, idx, grad_points, c, n, m, total);
  count_launch();
  return finish_launch();
}

extern "C" int gb_group_xyz(const float *xyz, const float *new_xyz, const int *idx, const float *rot, float *out, int b, int n, int m,
                            int nsample, float scale, int use_scale, long long out_scene_stride, gb_stream_t stream) {
  return group_xyz_impl(xyz, new_xyz, idx, rot, out, b, n, m, nsample, scale, use_scale, out_scene_stride, stream);
}

extern "C" int gb_group_xyz_feat(const float *xyz, const float *new_xyz, const int *idx, float *out_xyz, long long xyz_scene_stride,
                                 float scale, int use_scale, const float *points, float *out_feat, long long feat_scene_stride, int b,
                                 int c, int n, int npoints, int nsample, gb_stream_t stream) {
  if (b < 0 || c < 0 || n <= 0 || npoints < 0 || nsample < 0) return (int)cudaErrorInvalidValue;
  const size_t per = (size_t)npoints * nsample;
  if (b == 0 || per == 0) return 0;
  if (!xyz || !new_xyz || !idx || !out_xyz) return (int)cudaErrorInvalidValue;
  if (xyz_scene_stride < (long long)(3 * per)) return (int)cudaErrorInvalidValue;
  if (c == 0) return group_xyz_impl(xyz, new_xyz, idx, nullptr, out_xyz, b, n, npoints, nsample, scale, use_scale, xyz_scene_stride, stream);
  XyzFold fold;
  fold.xyz = xyz, fold.new_xyz = new_xyz, fold.out = out_xyz;
  fold.m = npoints, fold.ns = nsample, fold.scale = scale, fold.use_scale = use_scale;
  fold.out_stride = (size_t)xyz_scene_stride;
  fold.total = (size_t)b * npoints * (size_t)(nsample / 4);
  return group_fwd_impl(points, idx, out_feat, b, c, n, npoints, nsample, feat_scene_stride, stream, &fold);
}
