// query.cu -- ball query and cylinder query: shared-memory-tiled candidate scan, one WARP per query, warp-ballot
// ordered compaction, per-query early exit.
//
// Replaces query_ball_point_kernel (PointNet/_ext_src/src/ball_query_gpu.cu:9-44), ball_query_kernel_fast
// (pointnet2_batch/src/ball_query_gpu.cu:10-42) and query_cylinder_point_kernel (cylinder_query_gpu.cu:20-78).
// The reference gives each query ONE THREAD that walks the whole cloud from global/L2 (variant A even uses a single
// block per scene).  Here a CTA of 8 warps stages the cloud tile by tile in shared memory with the TMA engine
// (cp.async.bulk, double buffered, overlapping the scan of the previous tile), each lane tests one candidate per step,
// __ballot_sync + __popc give the hits their slot in index order, and a query stops as soon as it has nsample hits.
//
// Semantics kept bit for bit: candidates in ascending index; hit iff d2 < radius*radius (fp32, strict),
// for the cylinder additionally hmin < x_rot < hmax with (x_rot,y_rot,z_rot) = (p - q)^T R; the first hit pre-fills all
// nsample slots; no hit leaves zeros.  Floating-point contraction is the one nvcc applies to the reference source.
//
// Large clouds (n >= 4096) first go through a UNIFORM CELL GRID built per call (grid_build_kernel: bounding box,
// counting sort of the points by cell, all in one CTA's shared memory per scene).  A query then tests only the points of
// the cells its search region can reach (a conservative box, see grid_query_kernel) -- in any order -- and marks hits in
// a per-warp n-bit BITMAP in shared memory; reading the bitmap back in word order yields the first nsample hits in
// ascending index, exactly the reference's sequential scan.  The per-candidate arithmetic is the same instruction
// sequence as the full scan, and culling only removes points that cannot pass it, so results stay bit-exact.  When the
// grid would not cull (radius comparable to the scene) the build kernel says so and the full-scan kernel runs instead.
//
// Multi mode (gb_cylinder_query_multi, gb_cylinder_query_multi_radius): the nested cylinders of a grasp crop -- up to four
// depths (hmax) x four radii around the same seed, axis and hmin -- share ONE scan with the largest radius and depth; every
// hit is then classified by radius and depth and the up to 16 index lists of the seed are compacted from one hit buffer.
#include <float.h>

#include "common.cuh"

namespace gb {

__device__ __forceinline__ bool ball_hit(float qx, float qy, float qz, float x, float y, float z, float radius2) {
  return sqdist3(qx - x, qy - y, qz - z) < radius2;
}
// x_rot and the squared radial distance alone, the same instruction sequences as in cyl_hit (the multi read-out
// re-derives them per hit)
__device__ __forceinline__ void cyl_coords(const float (&r)[9], float qx, float qy, float qz, float x, float y, float z, float &xr,
                                           float &d2) {
  const float dx = x - qx, dy = y - qy, dz = z - qz;
  xr = __fmaf_rn(r[6], dz, __fmaf_rn(r[0], dx, __fmul_rn(r[3], dy)));
  const float yr = __fmaf_rn(r[7], dz, __fmaf_rn(r[1], dx, __fmul_rn(r[4], dy)));
  const float zr = __fmaf_rn(r[8], dz, __fmaf_rn(r[2], dx, __fmul_rn(r[5], dy)));
  d2 = __fmaf_rn(yr, yr, __fmul_rn(zr, zr));
}

// (x_rot, y_rot, z_rot) = (p - q)^T R with the reference's contraction (cylinder_query_gpu.cu:58-66, SASS-checked)
__device__ __forceinline__ bool cyl_hit(const float (&r)[9], float qx, float qy, float qz, float x, float y, float z, float radius2,
                                        float hmin, float hmax) {
  float xr, d2;
  cyl_coords(r, qx, qy, qz, x, y, z, xr, d2);
  return (d2 < radius2) && (xr > hmin) && (xr < hmax);
}

// Multi-depth cylinder query (the 4-depth loop of GraspWidthGrouping, TrainModel/modules.py:104-113): the cylinders of one
// call share axis, radius and hmin and differ in hmax only, so they are nested: ONE scan with the largest hmax finds every
// candidate, and a hit belongs to depth d iff x_rot < hmax[d].  idx is laid out [b, m, nd, nsample].
// Multi-radius on top (the four GraspWidthGrouping modules of GraspPoseStage2_seed_features_multi_scale.forward,
// TrainModel/graspbalance.py:104-107, differ in the cylinder radius only): a hit belongs to radius k iff d2 < r2[k], so
// the scan with the largest radius serves all nr x nd lists.  idx is laid out [nr, b, m, nd, nsample].
constexpr int kMaxDepths = 4;
struct HMax4 {
  float v[kMaxDepths];            // hmax of each depth
  float r2[kMaxDepths];           // radius * radius (fp32 product, as the single-radius launcher computes it) of each radius
  int nr;                         // number of radii (>= 1)
  unsigned long long rstride;     // elements of idx between two radii
};

// ---- uniform cell grid ------------------------------------------------------------------------------------------------
constexpr int kGridMaxCells = 4096;
constexpr int kGridMaxDim = 32;
constexpr int kGridThreads = 1024;
constexpr int kGridQueryWarps = 8;
constexpr int kMultiHitCap = 1024;  // records of the per-warp hit buffer of the multi read-out (>= 32 words x 32 bits)

struct __align__(16) GridHeader {
  float ox, oy, oz, inv;      // cell = clamp(floor((p - o) * inv))
  int gx, gy, gz, use_grid;   // use_grid = 0: the grid would not cull, run the full scan instead
  float maxabs, pad0, pad1, pad2;
};

// Monotone in x (fp32 subtract, multiply by a non-negative constant, floor, clamp), so every point with
// lo <= x <= hi lands in a cell between cell(lo) and cell(hi).  NaN -> cell 0, +-inf -> the border cells.
__device__ __forceinline__ int grid_cell(float x, float o, float inv, int g) {
  return min(g - 1, max(0, __float2int_rd(__fmul_rn(__fsub_rn(x, o), inv))));
}

// grid b, kGridThreads threads.  xyz [b,n,3] -> sorted [b,n] (x, y, z, original index as int bits) grouped by cell,
// cell_start [b, kGridMaxCells + 1], hdr [b].  reach = radius of a sphere around the query that contains the search region.
__global__ void __launch_bounds__(kGridThreads) grid_build_kernel(const float *__restrict__ xyz, int n, float reach, int force,
                                                                  float cell_frac,
                                                                  float4 *__restrict__ sorted, int *__restrict__ cell_start,
                                                                  GridHeader *__restrict__ hdr) {
  __shared__ int s_cnt[kGridMaxCells];
  __shared__ float s_red[32][6];
  __shared__ int s_wsum[32];
  __shared__ GridHeader s_h;
  const int scene = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  xyz += (size_t)scene * n * 3;
  sorted += (size_t)scene * n;
  cell_start += (size_t)scene * (kGridMaxCells + 1);

  // 1. bounding box of the finite coordinates
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = tid; i < n; i += kGridThreads) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = __ldg(xyz + 3 * (size_t)i + d);
      if (fabsf(v) <= FLT_MAX) lo[d] = fminf(lo[d], v), hi[d] = fmaxf(hi[d], v);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) s_red[warp][d] = lo[d], s_red[warp][3 + d] = hi[d];
  }
  for (int i = tid; i < kGridMaxCells; i += kGridThreads) s_cnt[i] = 0;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kGridThreads / 32; ++w) {
#pragma unroll
      for (int d = 0; d < 3; ++d) lo[d] = fminf(lo[d], s_red[w][d]), hi[d] = fmaxf(hi[d], s_red[w][3 + d]);
    }
    float ext[3], maxabs = 0.f, emax = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (lo[d] > hi[d]) lo[d] = hi[d] = 0.f;  // no finite coordinate at all
      ext[d] = hi[d] - lo[d];
      emax = fmaxf(emax, ext[d]);
      maxabs = fmaxf(maxabs, fmaxf(fabsf(lo[d]), fabsf(hi[d])));
    }
    GridHeader h;
    h.ox = lo[0], h.oy = lo[1], h.oz = lo[2];
    h.maxabs = maxabs, h.pad0 = h.pad1 = h.pad2 = 0.f;
    // cell edge = cell_frac x the reach of a query (its box then spans <= 2 / cell_frac + 1 cells per axis), >= extent / 32,
    // <= 4096 cells in total
    if (reach < 0.f) {  // nearest-neighbour search: twice the spacing n points have on a surface spanning the two largest extents
      const float e0 = fmaxf(ext[0], fmaxf(ext[1], ext[2])), e2 = fminf(ext[0], fminf(ext[1], ext[2]));
      const float e1 = ext[0] + ext[1] + ext[2] - e0 - e2;
      reach = 2.f * sqrtf(fmaxf(e0 * fmaxf(e1, 1e-3f * e0), 1e-30f) / (float)max(n, 1));
    }
    const float E = __fmaf_rn(reach, 1.0001f, 1e-5f * maxabs);
    float s = fmaxf(E * cell_frac, emax / kGridMaxDim);
    h.gx = h.gy = h.gz = 1, h.inv = 0.f;
    if (s > 0.f && s <= FLT_MAX) {
      for (;;) {
        h.gx = min(kGridMaxDim, (int)(ext[0] / s) + 1), h.gy = min(kGridMaxDim, (int)(ext[1] / s) + 1);
        h.gz = min(kGridMaxDim, (int)(ext[2] / s) + 1);
        if (h.gx * h.gy * h.gz <= kGridMaxCells) break;
        s *= 1.25f;
      }
      h.inv = 1.0f / s;
    }
    // share of the cloud's cells one query box touches; the grid pays off when that is small
    const float span = s > 0.f && s <= FLT_MAX ? 2.f * E / s + 1.f : 1.f;
    const float frac = fminf(1.f, span / h.gx) * fminf(1.f, span / h.gy) * fminf(1.f, span / h.gz);
    h.use_grid = (force || frac <= 0.3f) ? 1 : 0;
    s_h = h;
    hdr[scene] = h;
  }
  __syncthreads();
  const GridHeader h = s_h;
  if (!h.use_grid) return;
  const int ncells = h.gx * h.gy * h.gz;

  // 2. histogram
  for (int i = tid; i < n; i += kGridThreads) {
    const float x = __ldg(xyz + 3 * (size_t)i), y = __ldg(xyz + 3 * (size_t)i + 1), z = __ldg(xyz + 3 * (size_t)i + 2);
    const int c = (grid_cell(z, h.oz, h.inv, h.gz) * h.gy + grid_cell(y, h.oy, h.inv, h.gy)) * h.gx + grid_cell(x, h.ox, h.inv, h.gx);
    atomicAdd(&s_cnt[c], 1);
  }
  __syncthreads();
  // 3. exclusive scan over the cells (4 consecutive cells per thread)
  constexpr int kPer = kGridMaxCells / kGridThreads;
  int cnt[kPer], local = 0;
#pragma unroll
  for (int e = 0; e < kPer; ++e) {
    const int c = tid * kPer + e;
    cnt[e] = c < ncells ? s_cnt[c] : 0;
    local += cnt[e];
  }
  int incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = s_wsum[lane];
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += o;
    }
    s_wsum[lane] = wi - w;
  }
  __syncthreads();
  int run = s_wsum[warp] + incl - local;
#pragma unroll
  for (int e = 0; e < kPer; ++e) {
    const int c = tid * kPer + e;
    if (c < ncells) {
      s_cnt[c] = run;  // becomes the scatter cursor
      cell_start[c] = run;
    }
    run += cnt[e];
  }
  if (tid == 0) cell_start[ncells] = n;
  __syncthreads();
  // 4. scatter (order inside a cell is irrelevant: the bitmap restores index order)
  for (int i = tid; i < n; i += kGridThreads) {
    const float x = __ldg(xyz + 3 * (size_t)i), y = __ldg(xyz + 3 * (size_t)i + 1), z = __ldg(xyz + 3 * (size_t)i + 2);
    const int c = (grid_cell(z, h.oz, h.inv, h.gz) * h.gy + grid_cell(y, h.oy, h.inv, h.gy)) * h.gx + grid_cell(x, h.ox, h.inv, h.gx);
    const int pos = atomicAdd(&s_cnt[c], 1);
    sorted[pos] = make_float4(x, y, z, __int_as_float(i));
  }
}

// grid (ceil(m / 8), b), 8 warps, one WARP per query; dynamic smem 8 * words uint32 (words = ceil(n / 32)).
// Search region -> conservative box.  Ball: the cube of half-edge r around q.  Cylinder with R orthonormal to within err:
// p - q = R u with u = (x_rot, y_rot, z_rot), hmin < u0 < hmax, |(u1, u2)| < r, so along world axis i the offset lies in
// [min(R_i0 hmin, R_i0 hmax) - r |(R_i1, R_i2)|, max(R_i0 hmin, R_i0 hmax) + r |(R_i1, R_i2)|] (the cylinder's own
// bounding box, much tighter than its bounding sphere); a rotation further than 1e-3 from orthonormal searches the
// whole grid.  Every bound is widened by reach * (1e-4 + 4 err) + 1e-5 * (largest coordinate magnitude), orders of
// magnitude above the fp32 rounding of the test (a few ulp of the coordinates), so no point that passes it is culled.
// MULTI (cylinder only): hmax is the largest of the nd depths hm.v[]; the read-out classifies every hit by depth (x_rot
// re-derived from the original coordinates xyz_orig) and fills the nd index lists of the query, [nd, nsample] per query.
template <bool CYL, bool MULTI>
__global__ void __launch_bounds__(kGridQueryWarps * 32, MULTI ? 4 : 5) grid_query_kernel(const float *__restrict__ new_xyz, const float4 *__restrict__ sorted,
                                                                         const int *__restrict__ cell_start,
                                                                         const GridHeader *__restrict__ hdr, const float *__restrict__ rot,
                                                                         int *__restrict__ idx, int n, int m, float radius, float radius2,
                                                                         float reach, float hmin, float hmax, int nsample, int words,
                                                                         const float *__restrict__ xyz_orig, HMax4 hm, int nd) {
  extern __shared__ unsigned s_bm[];
  const int scene = blockIdx.y;
  const GridHeader h = hdr[scene];
  if (!h.use_grid) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned *bm = s_bm + (size_t)warp * words;
  for (int i = lane; i < words; i += 32) bm[i] = 0u;
  const int j = blockIdx.x * kGridQueryWarps + warp;
  if (j >= m) return;
  __syncwarp();
  sorted += (size_t)scene * n;
  cell_start += (size_t)scene * (kGridMaxCells + 1);
  const size_t qi = (size_t)scene * m + j;
  const float qx = __ldg(new_xyz + qi * 3), qy = __ldg(new_xyz + qi * 3 + 1), qz = __ldg(new_xyz + qi * 3 + 2);
  float r[9];
  float grow = 1e-4f;
  bool full = false;
  if (CYL) {
#pragma unroll
    for (int e = 0; e < 9; ++e) r[e] = __ldg(rot + qi * 9 + e);
    // columns of R are the axes the offsets are projected on
    const float c00 = r[0] * r[0] + r[3] * r[3] + r[6] * r[6], c11 = r[1] * r[1] + r[4] * r[4] + r[7] * r[7];
    const float c22 = r[2] * r[2] + r[5] * r[5] + r[8] * r[8], c01 = r[0] * r[1] + r[3] * r[4] + r[6] * r[7];
    const float c02 = r[0] * r[2] + r[3] * r[5] + r[6] * r[8], c12 = r[1] * r[2] + r[4] * r[5] + r[7] * r[8];
    const float err = fmaxf(fmaxf(fmaxf(fabsf(c00 - 1.f), fabsf(c11 - 1.f)), fmaxf(fabsf(c22 - 1.f), fabsf(c01))), fmaxf(fabsf(c02), fabsf(c12)));
    full = !(err < 1e-3f);  // also catches NaN
    grow += 4.f * err;
  }
  const float pad = __fmaf_rn(reach, grow, 1e-5f * fmaxf(fmaxf(h.maxabs, fabsf(qx)), fmaxf(fabsf(qy), fabsf(qz))));
  float blo[3], bhi[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (CYL) {
      const float a0 = r[3 * i] * hmin, a1 = r[3 * i] * hmax;
      const float rr = radius * sqrtf(r[3 * i + 1] * r[3 * i + 1] + r[3 * i + 2] * r[3 * i + 2]);
      blo[i] = fminf(a0, a1) - rr - pad, bhi[i] = fmaxf(a0, a1) + rr + pad;
    } else {
      blo[i] = -radius - pad, bhi[i] = radius + pad;
    }
  }
  int x0 = 0, x1 = h.gx - 1, y0 = 0, y1 = h.gy - 1, z0 = 0, z1 = h.gz - 1;
  if (!full && pad <= FLT_MAX) {
    x0 = grid_cell(qx + blo[0], h.ox, h.inv, h.gx), x1 = grid_cell(qx + bhi[0], h.ox, h.inv, h.gx);
    y0 = grid_cell(qy + blo[1], h.oy, h.inv, h.gy), y1 = grid_cell(qy + bhi[1], h.oy, h.inv, h.gy);
    z0 = grid_cell(qz + blo[2], h.oz, h.inv, h.gz), z1 = grid_cell(qz + bhi[2], h.oz, h.inv, h.gz);
  }

  // ---- candidates: the cells of a (z, y) row are contiguous in the sorted array.  The rows of the box (32 at a time, one
  // per lane: two independent cell_start loads each) are concatenated into one flat candidate range by a warp prefix sum, so
  // every lane has a candidate in every step whatever the row lengths; a lane finds the row of its flat index with a
  // cursor that only moves forward. ----
  const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
  for (int rb = 0; rb < nrows; rb += 32) {
    const int rr = rb + lane;
    int s0 = 0, cnt = 0;
    if (rr < nrows) {
      const int dz = rr / ny;
      const int row = ((z0 + dz) * h.gy + (y0 + rr - dz * ny)) * h.gx;
      s0 = __ldg(cell_start + row + x0);
      cnt = __ldg(cell_start + row + x1 + 1) - s0;
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int base = s0 - (incl - cnt);  // address of flat index f inside this lane's row = base + f
    int cur = 0;                         // row of this lane's current flat index
    for (int f0 = 0; f0 < total; f0 += 64) {
      float4 p[2];
      bool ok[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int f = f0 + 32 * u + lane;
        ok[u] = f < total;
        for (;;) {  // advance the cursor past the rows that end at or before f (empty rows included)
          const int inc = __shfl_sync(0xffffffffu, incl, cur);
          const bool adv = ok[u] && f >= inc;
          if (!__any_sync(0xffffffffu, adv)) break;
          cur += adv ? 1 : 0;
        }
        const int a = __shfl_sync(0xffffffffu, base, cur) + f;
        p[u] = ok[u] ? __ldg(sorted + a) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const bool hit = ok[u] && (CYL ? cyl_hit(r, qx, qy, qz, p[u].x, p[u].y, p[u].z, radius2, hmin, hmax)
                                       : ball_hit(qx, qy, qz, p[u].x, p[u].y, p[u].z, radius2));
        if (hit) {
          const int k = __float_as_int(p[u].w);
          atomicOr(&bm[k >> 5], 1u << (k & 31));
        }
      }
    }
  }
  __syncwarp();

  if (MULTI) {
    // ---- multi read-out.  Phase A: the set bits of the bitmap (hits of the scan with the largest radius and hmax) are
    // compacted, in ascending index, into a per-warp buffer of records  index | radius mask << 24 | depth mask << 28
    // (bit k of the radius mask: d2 < r2[k]; bit d of the depth mask: x_rot < hmax[d]) -- ONE warp prefix sum per 32
    // bitmap words whatever the number of lists.  Phase B (when the buffer is full, and at the end): every open list
    // (k, d) sweeps the buffer 32 records at a time and takes those with both of its bits, ballot-compacted.
    // Lane L keeps the count and the first hit of list L = k * nd + d. ----
    xyz_orig += (size_t)scene * n * 3;
    unsigned *hits = s_bm + (size_t)kGridQueryWarps * words + (size_t)warp * kMultiHitCap;
    const int nr = hm.nr, nlists = nr * nd;
    const unsigned lt_mask = (1u << lane) - 1u;
    int mycnt = lane < nlists ? 0 : nsample, myfirst = 0;
    int hcount = 0;
    auto flush = [&]() {
      __syncwarp();
      for (int L = 0; L < nlists; ++L) {
        int c = __shfl_sync(0xffffffffu, mycnt, L);
        if (c >= nsample) continue;
        const unsigned need = (0x01000000u << (L / nd)) | (0x10000000u << (L % nd));
        int *out = idx + (size_t)(L / nd) * hm.rstride + (qi * (size_t)nd + (L % nd)) * nsample;
        int first = __shfl_sync(0xffffffffu, myfirst, L);
        for (int base = 0; base < hcount && c < nsample; base += 32) {
          const unsigned rec = base + lane < hcount ? hits[base + lane] : 0u;
          const bool hit = (rec & need) == need;
          const unsigned mask = __ballot_sync(0xffffffffu, hit);
          if (!mask) continue;
          if (c == 0) first = (int)(__shfl_sync(0xffffffffu, rec, __ffs(mask) - 1) & 0xFFFFFFu);
          const int slot = c + __popc(mask & lt_mask);
          if (hit && slot < nsample) out[slot] = (int)(rec & 0xFFFFFFu);
          c += __popc(mask);
        }
        if (lane == L) mycnt = c, myfirst = first;
      }
      hcount = 0;
      __syncwarp();
    };
    for (int wb = 0; wb < words; wb += 32) {
      if (!__ballot_sync(0xffffffffu, mycnt < nsample)) break;  // every list is full
      const unsigned w = wb + lane < words ? bm[wb + lane] : 0u;
      if (!__ballot_sync(0xffffffffu, w != 0u)) continue;
      const int pc = __popc(w);
      int incl = pc;
#pragma unroll
      for (int dd = 1; dd < 32; dd <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, dd);
        if (lane >= dd) incl += o;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (hcount + total > kMultiHitCap) flush();  // total <= 1024 = kMultiHitCap
      int pos = hcount + incl - pc;
      for (unsigned t = w; t; t &= t - 1) {
        const int bit = __ffs(t) - 1;
        const unsigned k = (unsigned)(wb + lane) * 32u + bit;
        float xr, d2;
        cyl_coords(r, qx, qy, qz, __ldg(xyz_orig + 3 * (size_t)k), __ldg(xyz_orig + 3 * (size_t)k + 1), __ldg(xyz_orig + 3 * (size_t)k + 2), xr, d2);
        unsigned rec = k;
#pragma unroll
        for (int e = 0; e < kMaxDepths; ++e) {
          if (e < nr && d2 < hm.r2[e]) rec |= 0x01000000u << e;
          if (e < nd && xr < hm.v[e]) rec |= 0x10000000u << e;
        }
        hits[pos++] = rec;
      }
      hcount += total;
    }
    flush();
    for (int L = 0; L < nlists; ++L) {  // padding: the first hit (zeros when the list is empty)
      const int c = min(__shfl_sync(0xffffffffu, mycnt, L), nsample), f = __shfl_sync(0xffffffffu, myfirst, L);
      int *out = idx + (size_t)(L / nd) * hm.rstride + (qi * (size_t)nd + (L % nd)) * nsample;
      for (int sl = c + lane; sl < nsample; sl += 32) out[sl] = f;
    }
    return;
  }

  // ---- read the bitmap back in index order: first nsample hits, the rest padded with the first hit ----
  int *out = idx + qi * (size_t)nsample;
  int cnt = 0, first = 0;
  for (int wb = 0; wb < words && cnt < nsample; wb += 32) {
    unsigned w = wb + lane < words ? bm[wb + lane] : 0u;
    const unsigned any = __ballot_sync(0xffffffffu, w != 0u);
    if (!any) continue;
    const int pc = __popc(w);
    int incl = pc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (cnt == 0) {
      const int l0 = __ffs(any) - 1;
      const unsigned w0 = __shfl_sync(0xffffffffu, w, l0);
      first = (wb + l0) * 32 + __ffs(w0) - 1;
    }
    int pos = cnt + incl - pc;
    while (w && pos < nsample) {
      const int bit = __ffs(w) - 1;
      out[pos++] = (wb + lane) * 32 + bit;
      w &= w - 1;
    }
    cnt += total;
  }
  for (int sl = min(cnt, nsample) + lane; sl < nsample; sl += 32) out[sl] = first;
}

constexpr int kQueryWarps = 8;
constexpr int kQueryTile = 2016;  // points per shared-memory tile (23.6 KB, multiple of 32), two buffers fit the 48 KB static limit

// MULTI (CYL, QPW = kMaxDepths): the QPW slots of a warp are the nd depths of ONE query (hmax = hm.v[slot]) instead of
// QPW different queries; idx is [b, m, nd, nsample].
template <bool CYL, int QPW, bool MULTI = false>
__global__ void __launch_bounds__(kQueryWarps * 32) query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                                                                 const float *__restrict__ rot, int *__restrict__ idx, int n,
                                                                 int m, float radius2, float hmin, float hmax, int nsample,
                                                                 int use_bulk, const GridHeader *__restrict__ hdr, HMax4 hm, int nd) {
  __shared__ __align__(128) float tile[2][kQueryTile * 3];
  __shared__ uint64_t full[2];

  const int scene = blockIdx.y;
  if (hdr && hdr[scene].use_grid) return;  // this scene was answered by grid_query_kernel
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  xyz += (size_t)scene * n * 3;
  const int q0 = (blockIdx.x * kQueryWarps + warp) * (MULTI ? 1 : QPW);  // first query of this warp

  float qx[QPW], qy[QPW], qz[QPW];
  float r[CYL ? QPW : 1][9];
  float hmx[QPW];
  int cnt[QPW], first[QPW];
  int *out[QPW];
#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    const int j = MULTI ? q0 : q0 + q;
    const bool ok = j < m && (!MULTI || q < nd);
    const size_t qi = (size_t)scene * m + (j < m ? j : 0);
    hmx[q] = MULTI ? hm.v[q] : hmax;
    qx[q] = __ldg(new_xyz + qi * 3), qy[q] = __ldg(new_xyz + qi * 3 + 1), qz[q] = __ldg(new_xyz + qi * 3 + 2);
    if (CYL) {
#pragma unroll
      for (int e = 0; e < 9; ++e) r[q][e] = __ldg(rot + qi * 9 + e);
    }
    cnt[q] = ok ? 0 : nsample;  // out-of-range queries are "already full"
    first[q] = 0;
    out[q] = MULTI ? idx + (qi * (size_t)nd + (q < nd ? q : 0)) * nsample : idx + qi * (size_t)nsample;
  }

  const int ntiles = (n + kQueryTile - 1) / kQueryTile;
  if (use_bulk) {
    if (tid == 0) {
      mbar_init(&full[0], 1);
      mbar_init(&full[1], 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
      const int c0 = min(kQueryTile, n);
      mbar_arrive_expect_tx(&full[0], (uint32_t)c0 * 12u);
      bulk_g2s(tile[0], xyz, (uint32_t)c0 * 12u, &full[0]);
    }
  }

  const unsigned lt_mask = (1u << lane) - 1u;
  for (int t = 0; t < ntiles; ++t) {
    const int base_k = t * kQueryTile;
    const int tc = min(kQueryTile, n - base_k);
    const float *buf = tile[t & 1];
    if (use_bulk) {
      if (tid == 0 && t + 1 < ntiles) {
        const int nc = min(kQueryTile, n - (t + 1) * kQueryTile);
        mbar_arrive_expect_tx(&full[(t + 1) & 1], (uint32_t)nc * 12u);
        bulk_g2s(tile[(t + 1) & 1], xyz + (size_t)(t + 1) * kQueryTile * 3, (uint32_t)nc * 12u, &full[(t + 1) & 1]);
      }
      mbar_wait(&full[t & 1], (uint32_t)((t >> 1) & 1));
    } else {
      float *wbuf = tile[t & 1];
      for (int e = tid; e < tc * 3; e += kQueryWarps * 32) wbuf[e] = __ldg(xyz + (size_t)base_k * 3 + e);
      __syncthreads();
    }

    bool active = false;
#pragma unroll
    for (int q = 0; q < QPW; ++q) active |= cnt[q] < nsample;
    if (active) {
      for (int base = 0; base < tc; base += 32) {
        const int kk = base + lane;
        const bool valid = kk < tc;
        const int ks = valid ? kk : 0;
        const float x = buf[ks * 3], y = buf[ks * 3 + 1], z = buf[ks * 3 + 2];  // stride-3 words: conflict free
        bool any_left = false;
#pragma unroll
        for (int q = 0; q < QPW; ++q) {
          if (cnt[q] < nsample) {  // warp uniform
            const bool hit = valid && (CYL ? cyl_hit(r[q], qx[q], qy[q], qz[q], x, y, z, radius2, hmin, hmx[q])
                                           : ball_hit(qx[q], qy[q], qz[q], x, y, z, radius2));
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (mask) {
              if (cnt[q] == 0) first[q] = base_k + base + __ffs(mask) - 1;
              const int slot = cnt[q] + __popc(mask & lt_mask);
              if (hit && slot < nsample) out[q][slot] = base_k + kk;
              cnt[q] += __popc(mask);
            }
            any_left |= cnt[q] < nsample;
          }
        }
        if (!any_left) break;
      }
    }
    bool more = false;
#pragma unroll
    for (int q = 0; q < QPW; ++q) more |= cnt[q] < nsample;
    const int cta_more = __syncthreads_or(more ? 1 : 0);  // also fences reuse of the tile buffers
    if (!cta_more || t + 1 >= ntiles) {
      if (use_bulk && t + 1 < ntiles) mbar_wait(&full[(t + 1) & 1], (uint32_t)(((t + 1) >> 1) & 1));  // drain the in-flight copy
      break;
    }
  }

  // tail: slots [cnt, nsample) take the first hit; zeros when there was none (ball_query.cpp:24-26 relies on zeros)
#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    if (MULTI ? (q0 < m && q < nd) : (q0 + q < m)) {
      const int c = min(cnt[q], nsample);
      const int fill = c > 0 ? first[q] : 0;
      for (int s = c + lane; s < nsample; s += 32) out[q][s] = fill;
    }
  }
}

// nd = 0: one index list per query, [b, m, nsample].  nd >= 1 (cylinder only): nd nested depths hm.v[0..nd), hmax = the
// largest of them, idx [b, m, nd, nsample].
template <bool CYL, bool MULTI = false>
static int launch_query(const float *new_xyz, const float *xyz, const float *rot, int *idx, int b, int n, int m, float radius,
                        float hmin, float hmax, int nsample, cudaStream_t s, HMax4 hm = HMax4{{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, 1, 0ull},
                        int nd = 0) {
  if (b < 0 || n < 0 || m < 0 || nsample < 0) return (int)cudaErrorInvalidValue;
  if (b == 0 || m == 0 || nsample == 0) return 0;  // nothing to do (empty tensors have null data pointers)
  if (!idx) return (int)cudaErrorInvalidValue;
  if (n == 0)  // an empty cloud has no hit: every slot keeps the zero the reference's zero-filled output holds
    return (int)cudaMemsetAsync(idx, 0, (size_t)b * m * (nd > 0 ? nd : 1) * (MULTI ? hm.nr : 1) * nsample * sizeof(int), s);
  if (!new_xyz || !xyz || (CYL && !rot)) return (int)cudaErrorInvalidValue;
  if (b > 65535) {  // the batch rides on gridDim.y: larger batches go slab by slab
    if (MULTI) return (int)cudaErrorInvalidValue;  // the multi-radius layout [radii, b, ...] is not contiguous per slab
    for (int b0 = 0; b0 < b; b0 += 65535) {
      const int bb = b - b0 < 65535 ? b - b0 : 65535;
      const int rc = launch_query<CYL, MULTI>(new_xyz + (size_t)b0 * m * 3, xyz + (size_t)b0 * n * 3, rot ? rot + (size_t)b0 * m * 9 : nullptr,
                                              idx + (size_t)b0 * m * nsample, bb, n, m, radius, hmin, hmax, nsample, s, hm, nd);
      if (rc) return rc;
    }
    return 0;
  }
  const float radius2 = radius * radius;  // fp32 product, as ball_query_gpu.cu:22
  const int use_bulk = (n % 4 == 0) && (((uintptr_t)xyz & 15u) == 0);

  // ---- cell-grid path for large clouds; the build kernel decides per scene whether the grid culls enough ----
  GridHeader *hdr = nullptr;
  void *scratch = nullptr;
  const int words = (n + 31) / 32;
  // the grid pays for its build (one CTA per scene, ~30 us) from ~4M candidate tests per scene on (B200, 32 scenes:
  // n = m = 2048 113 us vs 165 us full scan; n = 2048, m = 1024 79 vs 62; n = m = 1024 85 vs 34)
  const bool grid_worth = n >= 4096 || (n >= 2048 && (long long)m * n >= (1LL << 22));
  if (g_tuning.query_mode != 1 && (grid_worth || g_tuning.query_mode == 2) && (size_t)words * kGridQueryWarps * 4 <= 96u * 1024u &&
      (!MULTI || n < (1 << 24))) {  // the multi read-out packs a point index in 24 bits
    const size_t sorted_bytes = (size_t)b * n * sizeof(float4);
    const size_t cells_bytes = (((size_t)b * (kGridMaxCells + 1) * sizeof(int)) + 15) & ~(size_t)15;
    cudaError_t e = scratch_alloc(&scratch, sorted_bytes + cells_bytes + (size_t)b * sizeof(GridHeader), s);
    if (e != cudaSuccess) return (int)e;
    float4 *sorted = reinterpret_cast<float4 *>(scratch);
    int *cell_start = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(scratch) + sorted_bytes);
    hdr = reinterpret_cast<GridHeader *>(reinterpret_cast<unsigned char *>(scratch) + sorted_bytes + cells_bytes);
    // radius of a sphere around the query that contains the search region (rounded up)
    double reach = fabs((double)radius);
    if (CYL) {
      const double hh = fmax(fabs((double)hmin), fabs((double)hmax));
      reach = sqrt(reach * reach + hh * hh);
    }
    const float reachf = (float)(reach * (1.0 + 1e-6));
    const float cell_frac = g_tuning.grid_cell_pct > 0 ? 0.01f * g_tuning.grid_cell_pct : 0.5f;
    grid_build_kernel<<<b, kGridThreads, 0, s>>>(xyz, n, reachf, g_tuning.query_mode == 2 ? 1 : 0, cell_frac, sorted, cell_start, hdr);
    count_launch();
    const size_t smem = ((size_t)words + (MULTI ? kMultiHitCap : 0)) * kGridQueryWarps * sizeof(unsigned);
    auto kern = grid_query_kernel<CYL, MULTI>;
    e = (cudaError_t)raise_smem_limit(kern, smem);
    if (e != cudaSuccess) {
      cudaFreeAsync(scratch, s);
      return (int)e;
    }
    kern<<<dim3((m + kGridQueryWarps - 1) / kGridQueryWarps, b), kGridQueryWarps * 32, smem, s>>>(new_xyz, sorted, cell_start, hdr, rot, idx, n, m,
                                                                                                 fabsf(radius), radius2, reachf, hmin, hmax, nsample,
                                                                                                 words, xyz, hm, nd);
    count_launch();
  }
  if (MULTI) {
    // scenes the grid declined (and small clouds): the full scan, the nd depths of a query as the slots of one warp, once
    // per radius
    dim3 grid((m + kQueryWarps - 1) / kQueryWarps, b);
    for (int kr = 0; kr < hm.nr; ++kr) {
      query_kernel<CYL, kMaxDepths, true><<<grid, kQueryWarps * 32, 0, s>>>(new_xyz, xyz, rot, idx + (size_t)kr * hm.rstride, n, m, hm.r2[kr],
                                                                             hmin, hmax, nsample, use_bulk, hdr, hm, nd);
      count_launch();
    }
    const int rc = finish_launch();
    if (scratch) cudaFreeAsync(scratch, s);
    return rc;
  }
  int qpw = g_tuning.query_qpw;
  if (qpw != 1 && qpw != 2 && qpw != 4) {
    const long warps = (long)b * m;
    qpw = warps >= 4L * 16 * num_sms() ? 4 : (warps >= 2L * 16 * num_sms() ? 2 : 1);
  }
  const int per_cta = kQueryWarps * qpw;
  dim3 grid((m + per_cta - 1) / per_cta, b);
  switch (qpw) {
    case 4: query_kernel<CYL, 4><<<grid, kQueryWarps * 32, 0, s>>>(new_xyz, xyz, rot, idx, n, m, radius2, hmin, hmax, nsample, use_bulk, hdr, hm, nd); break;
    case 2: query_kernel<CYL, 2><<<grid, kQueryWarps * 32, 0, s>>>(new_xyz, xyz, rot, idx, n, m, radius2, hmin, hmax, nsample, use_bulk, hdr, hm, nd); break;
    default: query_kernel<CYL, 1><<<grid, kQueryWarps * 32, 0, s>>>(new_xyz, xyz, rot, idx, n, m, radius2, hmin, hmax, nsample, use_bulk, hdr, hm, nd); break;
  }
  count_launch();
  const int rc = finish_launch();
  if (scratch) cudaFreeAsync(scratch, s);
  return rc;
}


// ---- three nearest neighbours through the cell grid -----------------------------------------------------------------------
// three_nn_kernel (three_nn.cu) tests every unknown point against all m known points (20.5 M tests per 20k-point scene).
// Here the known points are sorted into the uniform grid above (cell edge ~ the spacing of m points on a surface) and a
// thread scans the box of (2R+1)^3 cells around its point, R = 1, 2, 4, ... until its third-best squared distance is
// strictly below the squared distance to the nearest face of the box that still has cells behind it (minus a margin far
// above fp32 rounding) -- then no point outside the box can enter the result or tie with it.  Candidates arrive in cell
// order, so insertion compares (distance, index) lexicographically: exactly the result of the reference's ascending scan
// with strict `<` (interpolate_gpu.cu:38-56), bit for bit (same sqdist3 arithmetic).  NaN coordinates never win (d < b is
// false), as in the full scan.
constexpr int kNNGridThreads = 256;

template <bool WEIGHTS>
__global__ void __launch_bounds__(kNNGridThreads) three_nn_grid_kernel(const float *__restrict__ unknown, const float4 *__restrict__ sorted,
                                                                       const int *__restrict__ cell_start,
                                                                       const GridHeader *__restrict__ hdr, float *__restrict__ dist2,
                                                                       int *__restrict__ idx, float *__restrict__ weight, int n, int m) {
  const int scene = blockIdx.y;
  const int j = blockIdx.x * kNNGridThreads + threadIdx.x;
  if (j >= n) return;
  const GridHeader h = hdr[scene];
  sorted += (size_t)scene * m;
  cell_start += (size_t)scene * (kGridMaxCells + 1);
  const size_t uj = (size_t)scene * n + j;
  const float ux = __ldg(unknown + uj * 3), uy = __ldg(unknown + uj * 3 + 1), uz = __ldg(unknown + uj * 3 + 2);
  const float inf = __int_as_float(0x7f800000);
  const float s = h.inv > 0.f ? 1.0f / h.inv : inf;  // cell edge
  const int cx = grid_cell(ux, h.ox, h.inv, h.gx), cy = grid_cell(uy, h.oy, h.inv, h.gy), cz = grid_cell(uz, h.oz, h.inv, h.gz);
  float b1, b2, b3;
  int i1, i2, i3;
  for (int R = 1;; R <<= 1) {
    b1 = b2 = b3 = inf;
    i1 = i2 = i3 = 0;
    const int x0 = max(cx - R, 0), x1 = min(cx + R, h.gx - 1), y0 = max(cy - R, 0), y1 = min(cy + R, h.gy - 1);
    const int z0 = max(cz - R, 0), z1 = min(cz + R, h.gz - 1);
    for (int z = z0; z <= z1; ++z) {
      for (int y = y0; y <= y1; ++y) {
        const int row = (z * h.gy + y) * h.gx;
        const int e1 = __ldg(cell_start + row + x1 + 1);
        for (int e = __ldg(cell_start + row + x0); e < e1; ++e) {
          const float4 p = __ldg(sorted + e);
          const float d = sqdist3(ux - p.x, uy - p.y, uz - p.z);
          const int k = __float_as_int(p.w);
          if (d < b3 || (d == b3 && k < i3)) {  // (distance, index) ascending: what the reference's index-order scan keeps
            if (d < b1 || (d == b1 && k < i1)) {
              b3 = b2, i3 = i2, b2 = b1, i2 = i1, b1 = d, i1 = k;
            } else if (d < b2 || (d == b2 && k < i2)) {
              b3 = b2, i3 = i2, b2 = d, i2 = k;
            } else {
              b3 = d, i3 = k;
            }
          }
        }
      }
    }
    // distance to the nearest face of the box with cells behind it
    float dmin = inf;
    if (x0 > 0) dmin = fminf(dmin, ux - (h.ox + (float)x0 * s));
    if (x1 < h.gx - 1) dmin = fminf(dmin, (h.ox + (float)(x1 + 1) * s) - ux);
    if (y0 > 0) dmin = fminf(dmin, uy - (h.oy + (float)y0 * s));
    if (y1 < h.gy - 1) dmin = fminf(dmin, (h.oy + (float)(y1 + 1) * s) - uy);
    if (z0 > 0) dmin = fminf(dmin, uz - (h.oz + (float)z0 * s));
    if (z1 < h.gz - 1) dmin = fminf(dmin, (h.oz + (float)(z1 + 1) * s) - uz);
    if (dmin == inf) break;  // the box is the whole grid
    const float safe = dmin * 0.999f - 1e-4f * (h.maxabs + s);  // cell assignment and face positions are rounded in fp32
    if (safe > 0.f && b3 < safe * safe) break;
    if (!(ux == ux && uy == uy && uz == uz)) R = max(R, kGridMaxDim);  // a NaN query never terminates early: one pass over everything
  }
  // fewer than three finite distances: the reference leaves (inf, index 0) in the unfilled slots; an unfilled slot here holds
  // (inf, 0) as well because no candidate ever compared below inf.  But a known point whose distance is NaN is skipped by
  // both scans, and one at distance +inf ties with the initial state: the reference keeps index 0 (strict <), so do we.
  if (WEIGHTS) {
    b1 = __fsqrt_rn(b1), b2 = __fsqrt_rn(b2), b3 = __fsqrt_rn(b3);
    const float r1 = __frcp_rn(__fadd_rn(b1, 1e-8f)), r2 = __frcp_rn(__fadd_rn(b2, 1e-8f)), r3 = __frcp_rn(__fadd_rn(b3, 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r1, r3), r2);  // torch.sum over a 3-element inner dimension: (r0 + r2) + r1
    weight[uj * 3] = __fdiv_rn(r1, norm), weight[uj * 3 + 1] = __fdiv_rn(r2, norm), weight[uj * 3 + 2] = __fdiv_rn(r3, norm);
  }
  dist2[uj * 3] = b1, dist2[uj * 3 + 1] = b2, dist2[uj * 3 + 2] = b3;
  idx[uj * 3] = i1, idx[uj * 3 + 1] = i2, idx[uj * 3 + 2] = i3;
}

// Grid path of gb_three_nn / gb_three_nn_weights: worth its build (one CTA per scene) when many unknowns meet many knowns.
bool three_nn_grid_worth(int b, int n, int m) {
  if (g_tuning.query_mode == 1) return false;
  return m >= 256 && m < (1 << 24) && (long long)n * m >= (1LL << 21) && b <= 65535;
}

int three_nn_grid(const float *unknown, const float *known, float *dist2, int *idx, float *weight, int b, int n, int m, cudaStream_t s) {
  void *scratch = nullptr;
  const size_t sorted_bytes = (size_t)b * m * sizeof(float4);
  const size_t cells_bytes = (((size_t)b * (kGridMaxCells + 1) * sizeof(int)) + 15) & ~(size_t)15;
  cudaError_t e = scratch_alloc(&scratch, sorted_bytes + cells_bytes + (size_t)b * sizeof(GridHeader), s);
  if (e != cudaSuccess) return (int)e;
  float4 *sorted = reinterpret_cast<float4 *>(scratch);
  int *cell_start = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(scratch) + sorted_bytes);
  GridHeader *hdr = reinterpret_cast<GridHeader *>(reinterpret_cast<unsigned char *>(scratch) + sorted_bytes + cells_bytes);
  grid_build_kernel<<<b, kGridThreads, 0, s>>>(known, m, -1.f, 1, 1.0f, sorted, cell_start, hdr);
  count_launch();
  const dim3 grid((n + kNNGridThreads - 1) / kNNGridThreads, b);
  if (weight) three_nn_grid_kernel<true><<<grid, kNNGridThreads, 0, s>>>(unknown, sorted, cell_start, hdr, dist2, idx, weight, n, m);
  else three_nn_grid_kernel<false><<<grid, kNNGridThreads, 0, s>>>(unknown, sorted, cell_start, hdr, dist2, idx, nullptr, n, m);
  count_launch();
  const int rc = finish_launch();
  cudaFreeAsync(scratch, s);
  return rc;
}

}  // namespace gb

extern "C" int gb_ball_query(const float *new_xyz, const float *xyz, int *idx, int b, int n, int m, float radius, int nsample,
                             gb_stream_t stream) {
  return gb::launch_query<false>(new_xyz, xyz, nullptr, idx, b, n, m, radius, 0.f, 0.f, nsample, (cudaStream_t)stream);
}

extern "C" int gb_cylinder_query(const float *new_xyz, const float *xyz, const float *rot, int *idx, int b, int n, int m,
                                 float radius, float hmin, float hmax, int nsample, gb_stream_t stream) {
  return gb::launch_query<true>(new_xyz, xyz, rot, idx, b, n, m, radius, hmin, hmax, nsample, (cudaStream_t)stream);
}

/* Multi-depth / multi-radius cylinder query.  The loop over hmax_list of GraspWidthGrouping.forward (TrainModel/modules.py:
 * 104-113) calls cylinder_query once per depth, and GraspPoseStage2_seed_features_multi_scale.forward (TrainModel/
 * graspbalance.py:104-107) calls four such modules that differ in the radius only: nradii x ndepth (1..4 each) nested
 * cylinders per seed, ONE scan.  idx [nradii, b, m, ndepth, nsample]: idx[k, :, :, d, :] is bit-identical to
 * gb_cylinder_query(..., radii[k], hmin, hmax[d], ...).  radii / hmax are HOST arrays. */
extern "C" int gb_cylinder_query_multi_radius(const float *new_xyz, const float *xyz, const float *rot, int *idx, int b, int n, int m,
                                              const float *radii, int nradii, float hmin, const float *hmax, int ndepth, int nsample,
                                              gb_stream_t stream) {
  if (!hmax || !radii || ndepth < 1 || ndepth > gb::kMaxDepths || nradii < 1 || nradii > gb::kMaxDepths) return (int)cudaErrorInvalidValue;
  gb::HMax4 hm;
  float top = hmax[0];   // the largest non-NaN depth bounds the scan; a NaN depth matches nothing (x_rot < NaN is false)
  float rtop = 0.f;      // the radius with the largest fp32 square bounds it radially (a NaN radius matches nothing)
  for (int d = 0; d < gb::kMaxDepths; ++d) {
    hm.v[d] = hmax[d < ndepth ? d : 0];
    if (d < ndepth && hmax[d] == hmax[d] && (top != top || hmax[d] > top)) top = hmax[d];
    const float rk = radii[d < nradii ? d : 0];
    hm.r2[d] = rk * rk;  // fp32 product, as the single-radius launcher computes radius2
    if (d < nradii && rk == rk && fabsf(rk) > fabsf(rtop)) rtop = rk;
  }
  hm.nr = nradii;
  hm.rstride = (unsigned long long)b * m * ndepth * nsample;
  return gb::launch_query<true, true>(new_xyz, xyz, rot, idx, b, n, m, rtop, hmin, top, nsample, (cudaStream_t)stream, hm, ndepth);
}

extern "C" int gb_cylinder_query_multi(const float *new_xyz, const float *xyz, const float *rot, int *idx, int b, int n, int m,
                                       float radius, float hmin, const float *hmax, int ndepth, int nsample, gb_stream_t stream) {
  return gb_cylinder_query_multi_radius(new_xyz, xyz, rot, idx, b, n, m, &radius, 1, hmin, hmax, ndepth, nsample, stream);
}
