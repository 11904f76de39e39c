// scatter_private.cu -- group backward with WARP-PRIVATE accumulators (no sort, no atomics, no block barrier).
//
// Replaces group_points_grad_kernel (PointNet/_ext_src/src/group_points_gpu.cu:69-90) and group_points_grad_kernel_fast
// (pointnet2_batch/src/group_points_gpu.cu:9-22) for the shapes of the backbone (targets n <= ~2400, nsample a power of
// two >= 8 or a multiple of 32):   grad[b,c,idx[b,j,k]] += gout[b,c,j,k].
//
// The reference needs atomics because threads of different channels AND positions share one output row.  Here a WARP owns
// four channels of one scene for the whole kernel and keeps their sums for ALL n targets in its own slice of shared memory,
// acc[t] = float4 over the four channels (16 bytes per target: n = 2048 -> 32 KB per warp, 7 warps per SM).  It walks the
// scene's entries in storage order; lane L takes position L of a neighbourhood row, so
//   * every global load is a fully coalesced 128-byte (or 256-byte, nsample >= 64) piece of one channel row, read exactly
//     once, straight into registers (no staging in shared memory, no fill wavefronts);
//   * one step is one read-modify-write of 16 bytes per lane (LDS.128, 4 FADD, STS.128).  The 32 targets of a step come
//     from ONE neighbourhood row, whose indices are distinct (a ball / cylinder / kNN query lists a point once; the only
//     repeats are the padding copies of the first hit), so lanes never collide and nobody else touches the warp's slice:
//     plain stores are exact.  Rows are checked on the fly (strictly ascending = distinct: one shuffle and a vote); rows
//     that are not ascending (padding, kNN order, arbitrary caller indices) go through a match.any pass that sums
//     duplicates in registers first, so any index tensor gives the exact sum.
// The summation order is fixed (storage order per target), so the result is bit-reproducible run to run.
// Short rows: nsample 16 with n >= 512 packs four rows into a unit of 64 positions (8 lanes per row), updated one row after the
// other; nsample 8 (and 16 with few targets) lets a row fill a quarter / half of the warp, which then owns 16 / 8 channels as
// 4 / 2 planes of float4.
//
// Loads run R units ahead in registers (the only latency hiding a 7..14-warp SM has), as two alternating register sets --
// ptxas tracks all of this kernel's global loads with one scoreboard, see the main loop.  HBM traffic is the algorithmic
// minimum: gout once, grad once; idx is re-read from L2 once per channel group.
#include "common.cuh"

namespace gb {

constexpr size_t kPrivSmemBudget = 227u * 1024u;
#ifndef GB_PRIV_RING_NUM
#define GB_PRIV_RING_NUM 3  // two-channel units: ring depth x 3/2
#endif
#ifndef GB_PRIV_TWOSET
#define GB_PRIV_TWOSET -1  // experiments: 0 = slot ring, 1 = two sets refilled in one burst, 2 = two sets, spread refill -- for every shape
#endif


template <int CW, int VL>
struct PrivUnit {
  int t[VL];
  float v[CW][VL];
};

template <int VL>
struct PrivVec;
template <>
struct PrivVec<1> {
  static __device__ __forceinline__ void ldf(float (&d)[1], const float *p) {
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(d[0]) : "l"(p));
  }
  static __device__ __forceinline__ void ldi(int (&d)[1], const int *p) { d[0] = __ldg(p); }
};
template <>
struct PrivVec<2> {
  static __device__ __forceinline__ void ldf(float (&d)[2], const float *p) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(d[0]), "=f"(d[1]) : "l"(p));
  }
  static __device__ __forceinline__ void ldi(int (&d)[2], const int *p) {
    const int2 t = __ldg(reinterpret_cast<const int2 *>(p));
    d[0] = t.x, d[1] = t.y;
  }
};
template <>
struct PrivVec<4> {
  static __device__ __forceinline__ void ldf(float (&d)[4], const float *p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "l"(p));
  }
  static __device__ __forceinline__ void ldi(int (&d)[4], const int *p) {
    const int4 t = __ldg(reinterpret_cast<const int4 *>(p));
    d[0] = t.x, d[1] = t.y, d[2] = t.z, d[3] = t.w;
  }
};

// CW floats (2 or 4) of one target: the unit of a read-modify-write
template <int CW>
struct PrivAcc;
template <>
struct PrivAcc<4> {
  typedef float4 T;
  static __device__ __forceinline__ void add(float4 &a, const float (&x)[4]) { a.x += x[0], a.y += x[1], a.z += x[2], a.w += x[3]; }
  static __device__ __forceinline__ float get(const float4 &a, int j) { return j == 0 ? a.x : (j == 1 ? a.y : (j == 2 ? a.z : a.w)); }
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <>
struct PrivAcc<2> {
  typedef float2 T;
  static __device__ __forceinline__ void add(float2 &a, const float (&x)[2]) { a.x += x[0], a.y += x[1]; }
  static __device__ __forceinline__ float get(const float2 &a, int j) { return j == 0 ? a.x : a.y; }
  static __device__ __forceinline__ float2 zero() { return make_float2(0.f, 0.f); }
};

// Padded rows (an ascending run followed by copies of the row's first target t0: how ball_query_gpu.cu:31-40 and
// cylinder_query_gpu.cu:68-75 fill rows with fewer than nsample hits): every lane hands over the sum of its copies' values;
// the LG lanes of a row (and plane) add them up and the row's first lane updates t0 once.  Two forms: inline (a call waits
// for every scoreboard, i.e. drains all loads in flight: the kernels with many padded rows) and out of line behind scalar
// arguments (keeps the unrolled hot loop small: everything else).
template <int CW, int NRG, int LG>
__device__ __forceinline__ void priv_add_copies(typename PrivAcc<CW>::T *acc, int t0, float s0, float s1, float s2, float s3) {
  float sum[CW];
  sum[0] = s0, sum[1] = s1;
  if (CW == 4) sum[CW - 2] = s2, sum[CW - 1] = s3;
#pragma unroll
  for (int d = LG / 2; d; d >>= 1) {
#pragma unroll
    for (int j = 0; j < CW; ++j) sum[j] += __shfl_xor_sync(0xffffffffu, sum[j], d);
  }
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int r = 0; r < NRG; ++r) {  // the rows of a unit one after the other: two of them may share their first target
    if ((NRG == 1 || lane / LG == r) && (lane & (LG - 1)) == 0) {
      typename PrivAcc<CW>::T a = acc[t0];
      PrivAcc<CW>::add(a, sum);
      acc[t0] = a;
    }
    __syncwarp();
  }
}
template <int CW, int NRG, int LG>
__device__ __noinline__ void priv_add_copies_call(typename PrivAcc<CW>::T *acc, int t0, float s0, float s1, float s2, float s3) {
  priv_add_copies<CW, NRG, LG>(acc, t0, s0, s1, s2, s3);
}

// Rows that are neither ascending nor padded (kNN order, arbitrary caller tensors, out-of-range entries), out of line: per
// step, repeats among the active lanes' targets (match.any) are summed in registers first, then every target is updated
// once.  acc = this lane's plane; the unit is read from local memory (the caller's copy).
template <int CW, int VL, int NRG, int LG>
__device__ __noinline__ void priv_slow_unit(typename PrivAcc<CW>::T *acc, int n, int h, const PrivUnit<CW, VL> *up) {
  typedef typename PrivAcc<CW>::T AccT;
  const int lane = threadIdx.x & 31;
  const int rg = lane / LG;
  const PrivUnit<CW, VL> u = *up;
#pragma unroll 1
  for (int r = 0; r < NRG; ++r) {
#pragma unroll 1
    for (int q = 0; q < VL; ++q) {
      const int t = u.t[q];
      float x[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) x[j] = u.v[j][q];
      bool valid = (NRG == 1 || rg == r) && (unsigned)t < (unsigned)n;
      const unsigned key = valid ? ((unsigned)t | ((unsigned)h << 28)) : (0x80000000u | (unsigned)lane);
      const unsigned mask = __match_any_sync(0xffffffffu, key);
      unsigned dups = __ballot_sync(0xffffffffu, (mask & (mask - 1)) != 0u);
      while (dups) {
        const int leader = __ffs(dups) - 1;
        const unsigned gm = __shfl_sync(0xffffffffu, mask, leader);
        const bool in = (gm >> lane) & 1u;
        float s[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) s[j] = in ? x[j] : 0.f;
#pragma unroll
        for (int d = 16; d; d >>= 1) {
#pragma unroll
          for (int j = 0; j < CW; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], d);
        }
        if (lane == leader) {
#pragma unroll
          for (int j = 0; j < CW; ++j) x[j] = s[j];
        } else if (in) {
          valid = false;
        }
        dups &= ~gm;
      }
      if (valid) {
        AccT a = acc[t];
        PrivAcc<CW>::add(a, x);
        acc[t] = a;
      }
      __syncwarp();
    }
  }
}

// S lanes span a row piece (32, 16 or 8); H = 32 / S planes of CW channels (S < 32: the planes share the row).  VL consecutive
// positions per lane and load (LDG.32/64/128): a unit is 32 * VL positions (S == 32) = NRG whole rows (NRG = 1: one row or a
// piece of a longer one).  The NRG rows of a unit are updated one after the other (their targets may coincide), the VL
// positions of a lane together.  R units in flight; MAXT bounds the block (register budget).
// grid: ceil(tasks / W) blocks of W warps; task = (scene, group of CW * H channels).  dynamic smem: W * H * n * 4 CW bytes.
template <int CW, int S, int VL, int NRG, int R, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) scatter_private_kernel(const float *__restrict__ src, const int *__restrict__ idx,
                                                                  float *__restrict__ grad, int c, int n, int per, int groups,
                                                                  int tasks, int overwrite, size_t src_stride, int ns, int split,
                                                                  int dry) {
  constexpr int H = 32 / S, PU = S * VL, LG = S == 32 ? 32 / NRG : S;  // LG lanes hold one row (piece)
  // how the loads run ahead (see the main loop): two register sets for the kernels whose lanes span rows (S = 32), the slot
  // ring for the channel-plane kernels (short rows, few targets)
  constexpr bool kTwoSets = GB_PRIV_TWOSET >= 0 ? GB_PRIV_TWOSET != 0 : S == 32;
  constexpr bool kSpread = GB_PRIV_TWOSET != 1;  // refill spread over the first G - 2 units of the other set
  typedef typename PrivAcc<CW>::T AccT;
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // task = (scene, channel group); `split` consecutive warps of a block share a task and take every split-th unit of its
  // rows (small launches: twice or four times the warps per SM); their accumulators are added up at the write-out
  const int wtask = blockIdx.x * (blockDim.x >> 5) + warp;
  const int task = wtask / split, part = wtask - task * split;
  const bool live = task < tasks;
  if (!live && split == 1) return;  // unsplit: warps are independent, no block-level barrier anywhere below
  const int scene = live ? task / groups : 0, grp = live ? task - scene * groups : 0;
  const int k = lane & (S - 1), h = lane / S;
  const int rg = lane / LG;  // row of the unit this lane works on (0 when NRG == 1)
  const bool row_end = (lane & (LG - 1)) == LG - 1;
  AccT *acc_w = reinterpret_cast<AccT *>(s_raw) + (size_t)warp * H * n;  // [H][n] x CW floats
  for (int i = lane; i < H * n; i += 32) acc_w[i] = PrivAcc<CW>::zero();
  __syncwarp();
  AccT *acc = acc_w + (size_t)h * n;  // this lane's plane

  // channels past c (ragged last group) read the last valid row: their sums are never written out
  const int ch0 = grp * CW * H + CW * h;
  const int *ip = idx + (size_t)scene * per + VL * k;
  const float *g[CW];
#pragma unroll
  for (int j = 0; j < CW; ++j) g[j] = src + (size_t)scene * src_stride + (size_t)min(ch0 + j, c - 1) * per + VL * k;
  const int units = per / PU;

  auto load = [&](PrivUnit<CW, VL> &u, int unit) {  // units past the end re-read the last one (never processed)
    const int off = min(unit, units - 1) * PU;
    PrivVec<VL>::ldi(u.t, ip + off);
#pragma unroll
    for (int j = 0; j < CW; ++j) PrivVec<VL>::ldf(u.v[j], g[j] + off);
  };
  // Row classes (warp-uniform):
  //   0  strictly ascending along every row of the unit, all inside [0, n): the targets of a row step are distinct;
  //   1  an ascending run followed by copies of the row's first target (how ball_query_gpu.cu:31-40 and
  //      cylinder_query_gpu.cu:68-75 fill rows with fewer than nsample hits): the copies are summed in registers and
  //      added once.  Needs the row's first target, i.e. whole rows in a unit (nsample <= 32 * VL);
  //   2  anything else (kNN order, arbitrary caller tensors, out-of-range entries): the match.any path.
  // t0 = first target of the lane's row (valid for class 1 only).
  auto classify = [&](const PrivUnit<CW, VL> &u, int &t0) {
    const int next = __shfl_down_sync(0xffffffffu, u.t[0], 1);
    bool asc = (unsigned)u.t[0] < (unsigned)n && (unsigned)u.t[VL - 1] < (unsigned)n && (row_end || u.t[VL - 1] < next);
#pragma unroll
    for (int q = 1; q < VL; ++q) asc = asc && u.t[q - 1] < u.t[q];
    if (__all_sync(0xffffffffu, asc)) return 0;
    if (ns > PU) return 2;
    t0 = __shfl_sync(0xffffffffu, u.t[0], lane & ~(LG - 1));
    const bool first = (lane & (LG - 1)) == 0;  // this lane's first position opens its row
    bool ok = true;
#pragma unroll
    for (int q = 0; q < VL; ++q) {
      const int a = u.t[q], b = q + 1 < VL ? u.t[q + 1 < VL ? q + 1 : q] : next;
      const bool a_pad = a == t0 && !(first && q == 0);
      ok = ok && (unsigned)a < (unsigned)n;
      if (q + 1 < VL || !row_end) ok = ok && (a_pad ? b == t0 : (a < b || b == t0));
    }
    return __all_sync(0xffffffffu, ok) ? 1 : 2;
  };

  float dry_sum = 0.f;
  auto process = [&](const PrivUnit<CW, VL> &u) {
    if (dry) {  // experiment: the load pipeline alone (no shared-memory work)
#pragma unroll
      for (int q = 0; q < VL; ++q)
#pragma unroll
        for (int j = 0; j < CW; ++j) dry_sum += u.v[j][q] * (float)u.t[q];
      return;
    }
    int t0 = 0;
    const int cls = classify(u, t0);
    if (cls == 2) {
      const PrivUnit<CW, VL> copy = u;  // addressable copy: the ring itself stays in registers
      priv_slow_unit<CW, VL, NRG, LG>(acc, n, h, &copy);
      return;
    }
    if (cls == 0) {
#pragma unroll
      for (int r = 0; r < NRG; ++r) {
        if (NRG == 1 || rg == r) {
          AccT a[VL];
#pragma unroll
          for (int q = 0; q < VL; ++q) a[q] = acc[u.t[q]];
#pragma unroll
          for (int q = 0; q < VL; ++q) {
            float x[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) x[j] = u.v[j][q];
            PrivAcc<CW>::add(a[q], x);
            acc[u.t[q]] = a[q];
          }
        }
        __syncwarp();  // the next row may name the same targets on other lanes
      }
      return;
    }
    // class 1: the copies of t0 leave the row (summed and added once, out of line); what is left is strictly ascending
    const bool first = (lane & (LG - 1)) == 0;
    bool pad[VL];
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < VL; ++q) {
      pad[q] = u.t[q] == t0 && !(first && q == 0);
#pragma unroll
      for (int j = 0; j < CW; ++j) sum[j] += pad[q] ? u.v[j][q] : 0.f;
    }
    if (kTwoSets) priv_add_copies<CW, NRG, LG>(acc, t0, sum[0], sum[1], sum[2], sum[3]);
    else priv_add_copies_call<CW, NRG, LG>(acc, t0, sum[0], sum[1], sum[2], sum[3]);
#pragma unroll 1
    for (int r = 0; r < NRG; ++r) {
      if (NRG == 1 || rg == r) {
#pragma unroll
        for (int q = 0; q < VL; ++q) {
          if (!pad[q]) {
            AccT a = acc[u.t[q]];
            float x[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) x[j] = u.v[j][q];
            PrivAcc<CW>::add(a, x);
            acc[u.t[q]] = a;
          }
        }
      }
      __syncwarp();
    }
  };

  // Loads run ahead in registers (the only latency hiding a 7..14-warp SM has).  ptxas tracks every global load of this
  // kernel with ONE scoreboard, so the first use of a loaded register after the loop's back edge waits for ALL loads in
  // flight, the youngest included: a slot-by-slot ring stalls for a full memory latency once per round (14-17 % of a warp's
  // time in ncu's source view).  Hence two register sets of G units: set B is refilled while the first G - 2 units of set A
  // are processed (spread over them: one burst of all G units doubles the latency of the shared-memory reads behind it), so
  // whatever is in flight when B's wait comes is at least two units old; then the roles swap.
  const int my_units = live ? (units - part + split - 1) / split : 0;  // units part, part + split, ...
  constexpr int G = R / 2;
  PrivUnit<CW, VL> ring[kTwoSets ? 2 * G : R];
  if (kTwoSets) {
#pragma unroll
    for (int g = 0; g < G; ++g) load(ring[g], part + g * split);
    for (int u0 = 0; u0 < my_units; u0 += 2 * G) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (u0 + g < my_units) process(ring[g]);  // warp-uniform
#pragma unroll
        for (int q = 0; q < G; ++q)
          if ((kSpread ? q * (G - 2) / G : 0) == g) load(ring[G + q], part + (u0 + G + q) * split);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (u0 + G + g < my_units) process(ring[G + g]);
#pragma unroll
        for (int q = 0; q < G; ++q)
          if ((kSpread ? q * (G - 2) / G : 0) == g) load(ring[q], part + (u0 + 2 * G + q) * split);
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) load(ring[r], part + r * split);
    for (int u0 = 0; u0 < my_units; u0 += R) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (u0 + r < my_units) process(ring[r]);
        load(ring[r], part + (u0 + R + r) * split);
      }
    }
  }
  __syncwarp();
  if (split > 1) {
    __syncthreads();  // the partners' accumulators are complete
    if (!live || part != 0) return;
  }
  if (dry) acc[lane] = acc[lane], grad[(size_t)scene * c * n + lane] = dry_sum;

  // rows of the warp's channels, written once (coalesced along the targets)
#pragma unroll
  for (int hh = 0; hh < H; ++hh) {
    const int cbase = grp * CW * H + CW * hh;
    const AccT *a = acc_w + (size_t)hh * n;
    for (int t = lane; t < n; t += 32) {
      AccT x = a[t];
      for (int p = 1; p < split; ++p) {  // partners' slices lie behind this warp's, in part order: a fixed summation order
        const AccT y = a[(size_t)p * H * n + t];
        float ys[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) ys[j] = PrivAcc<CW>::get(y, j);
        PrivAcc<CW>::add(x, ys);
      }
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        if (cbase + j < c) {
          float *dst = grad + ((size_t)scene * c + cbase + j) * n + t;
          const float v = PrivAcc<CW>::get(x, j);
          *dst = overwrite ? v : *dst + v;
        }
      }
    }
  }
}

// lanes that span a row piece: 32 for nsample >= 32; short rows (nsample 8 / 16) either share the warp between 32 / nsample
// channel planes (S = nsample, one position per lane) or sit several to a unit like the long ones (S = 32, VL = 2, NRG rows)
// B200, 32 scenes, nsample 16: four rows per unit 86 / 88 / 50 us (n = 512, C = 256 / n = 1024, m = 512 / n = 512, m = 256)
// against 101 / 114 / 60 us with planes (26 warp instructions per 32 elements drop to ~18); n = 256: 49 against 45 us.
static int priv_lanes(int nsample, long long npoints, int n) {
  if (nsample >= 32) return 32;
  const int knob = g_tuning.priv_rows;
  const bool rows = knob == 2 || (knob == 0 && nsample == 16 && n >= 512);
  return rows && (npoints * nsample) % 64 == 0 ? 32 : nsample;  // whole units of 64 positions
}

// channels per warp and plane: 4 (float4 accumulators) while at least eight warps' worth fits in shared memory, else 2
static int priv_cw(int n, int H) {
  const int knob = g_tuning.priv_cw;
  if (knob == 2 || knob == 4) return knob;
  return (size_t)8 * H * n * 16 <= kPrivSmemBudget ? 4 : 2;
}

// shapes the private-accumulator path takes: a row piece per step (nsample 8, 16, or a multiple of 32), at least six warps
// of accumulators per SM, enough (scene, channel group) tasks to fill the GPU
bool scatter_private_supported(int b, int c, int n, int npoints, int nsample, size_t src_stride, const float *src, const int *idx) {
  if (g_tuning.scatter_mode & 8) return false;
  const int S = priv_lanes(nsample, npoints, n);
  if (nsample < 8 || (nsample & (nsample - 1)) != 0) return false;  // rows are pieces of a warp step: a power of two
  const int H = 32 / S, CW = priv_cw(n, H);
  if ((size_t)6 * H * n * 4 * CW > kPrivSmemBudget) return false;
  const size_t per = (size_t)npoints * nsample;
  if (per >= (1u << 30) || src_stride % 4 != 0 || per % 4 != 0) return false;
  if ((((uintptr_t)src | (uintptr_t)idx) & 15u) != 0) return false;
  const long long tasks = (long long)b * ((c + CW * H - 1) / (CW * H));
  const long long min_tasks = g_tuning.scatter_mode & 16 ? 1 : 2LL * num_sms();
  return tasks >= min_tasks && tasks < (1LL << 30);
}

template <int CW, int S, int VL, int NRG, int R, int MAXT>
static int launch_private(const float *src, size_t src_stride, const int *idx, float *grad, int c, int n, int per, int groups, int tasks,
                          int W, int split, int overwrite, int nsample, cudaStream_t s) {
  constexpr int H = 32 / S;
  const size_t smem = (size_t)W * H * n * 4 * CW;
  auto kern = scatter_private_kernel<CW, S, VL, NRG, R, MAXT>;
  if (int rc_ = raise_smem_limit(kern, smem)) return rc_;
  const long long warps = (long long)tasks * split;
  kern<<<(unsigned)((warps + W - 1) / W), W * 32, smem, s>>>(src, idx, grad, c, n, per, groups, tasks, overwrite, src_stride, nsample, split,
                                                            g_tuning.priv_dry);
  count_launch();
  return finish_launch();
}

template <int CW>
static int scatter_private_cw(const float *src, size_t src_stride, const int *idx, float *grad, int b, int c, int n, int npoints,
                              int nsample, int overwrite, cudaStream_t s) {
  const int S = priv_lanes(nsample, npoints, n), H = 32 / S;
  const int per = npoints * nsample;
  const int groups = (c + CW * H - 1) / (CW * H);
  const int tasks = b * groups;
  int wmax = (int)(kPrivSmemBudget / ((size_t)H * n * 4 * CW));
  wmax = wmax > 16 ? 16 : wmax;
  // warps per block: all that fit when the launch fills the GPU anyway, else spread the tasks over the SMs ...
  int W = (tasks + num_sms() - 1) / num_sms();
  W = W > wmax ? wmax : (W < 1 ? 1 : W);
  if (g_tuning.scatter_cc > 0 && g_tuning.scatter_cc <= wmax) W = g_tuning.scatter_cc;
  // ... and when fewer warps than fit would be busy, 2 or 4 warps share a task (every split-th unit of its rows each)
  int split = 1;
  while (split < 4 && W * split * 2 <= wmax && (long long)npoints * nsample / (32 * 2 * split) >= 64) split *= 2;
  if (g_tuning.priv_split == 1 || g_tuning.priv_split == 2 || g_tuning.priv_split == 4) {
    split = g_tuning.priv_split;
    while (split > 1 && W * split > wmax) split >>= 1;
  }
  W *= split;
  // positions per lane and load; a unit of 32 * VL positions holds NRG = 32 * VL / nsample whole rows (or a piece of one)
  int VL = 1;
  if (S == 32) {
    VL = nsample >= 64 ? 4 : 2;  // a unit of two rows for nsample 32 / 64 (B200 sweep: profiles/r02_bwd_private_sweep.json)
    const int knob = g_tuning.priv_vl;
    if ((knob == 1 || knob == 2 || knob == 4) && (nsample % (32 * knob) == 0 || (32 * knob) % nsample == 0) && (nsample >= 32 || knob > 1)) VL = knob;
    while (VL > 1 && per % (32 * VL) != 0) VL >>= 1;  // units are whole: an odd number of short rows takes narrower loads
  }
  const int NRG = S == 32 && nsample < 32 * VL ? 32 * VL / nsample : 1;
  // ring depths (units in flight per warp): as deep as the register budget of the block size allows; two-channel units are
  // half as large, so their rings are deeper
#define GB_PRIV(SV, VV, NV, R8, R16_)                                                                                                \
  return W <= 8 ? launch_private<CW, SV, VV, NV, R8, 256>(src, src_stride, idx, grad, c, n, per, groups, tasks, W, split, overwrite, nsample, s)    \
                : launch_private<CW, SV, VV, NV, (CW == 2 ? (R16_ * GB_PRIV_RING_NUM) / 2 : R16_), 512>(src, src_stride, idx, grad, c, n, per, groups, tasks, W, split, overwrite, nsample, s)
  if (S == 32 && VL == 4 && NRG == 8) GB_PRIV(32, 4, 8, 8, 4);
  if (S == 32 && VL == 4 && NRG == 4) GB_PRIV(32, 4, 4, 8, 4);
  if (S == 32 && VL == 4 && NRG == 2) GB_PRIV(32, 4, 2, 8, 4);
  if (S == 32 && VL == 4) GB_PRIV(32, 4, 1, 8, 4);
  if (S == 32 && VL == 2 && NRG == 8) GB_PRIV(32, 2, 8, 12, 8);
  if (S == 32 && VL == 2 && NRG == 4) GB_PRIV(32, 2, 4, 12, 8);
  if (S == 32 && VL == 2 && NRG == 2) GB_PRIV(32, 2, 2, 12, 8);
  if (S == 32 && VL == 2) GB_PRIV(32, 2, 1, 12, 8);
  if (S == 32) GB_PRIV(32, 1, 1, 16, 10);
  if (S == 16) GB_PRIV(16, 1, 1, 16, 10);
  GB_PRIV(8, 1, 1, 16, 10);
#undef GB_PRIV
}

int scatter_private(const float *src, size_t src_stride, const int *idx, float *grad, int b, int c, int n, int npoints, int nsample,
                    int overwrite, cudaStream_t s) {
  const int H = 32 / priv_lanes(nsample, npoints, n);
  if (priv_cw(n, H) == 2) return scatter_private_cw<2>(src, src_stride, idx, grad, b, c, n, npoints, nsample, overwrite, s);
  return scatter_private_cw<4>(src, src_stride, idx, grad, b, c, n, npoints, nsample, overwrite, s);
}

}  // namespace gb
