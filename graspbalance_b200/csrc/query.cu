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
#include "common.cuh"

namespace gb {

constexpr int kQueryWarps = 8;
constexpr int kQueryTile = 2016;  // points per shared-memory tile (23.6 KB, multiple of 32), two buffers fit the 48 KB static limit

template <bool CYL, int QPW>
__global__ void __launch_bounds__(kQueryWarps * 32) query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                                                                 const float *__restrict__ rot, int *__restrict__ idx, int n,
                                                                 int m, float radius2, float hmin, float hmax, int nsample,
                                                                 int use_bulk) {
  __shared__ __align__(128) float tile[2][kQueryTile * 3];
  __shared__ uint64_t full[2];

  const int scene = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  xyz += (size_t)scene * n * 3;
  const int q0 = (blockIdx.x * kQueryWarps + warp) * QPW;  // first query of this warp

  float qx[QPW], qy[QPW], qz[QPW];
  float r[CYL ? QPW : 1][9];
  int cnt[QPW], first[QPW];
  int *out[QPW];
#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    const int j = q0 + q;
    const bool ok = j < m;
    const size_t qi = (size_t)scene * m + (ok ? j : 0);
    qx[q] = __ldg(new_xyz + qi * 3), qy[q] = __ldg(new_xyz + qi * 3 + 1), qz[q] = __ldg(new_xyz + qi * 3 + 2);
    if (CYL) {
#pragma unroll
      for (int e = 0; e < 9; ++e) r[q][e] = __ldg(rot + qi * 9 + e);
    }
    cnt[q] = ok ? 0 : nsample;  // out-of-range queries are "already full"
    first[q] = 0;
    out[q] = idx + qi * (size_t)nsample;
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
            bool hit;
            if (CYL) {
              const float dx = x - qx[q], dy = y - qy[q], dz = z - qz[q];
              const float xr = __fmaf_rn(r[q][6], dz, __fmaf_rn(r[q][0], dx, __fmul_rn(r[q][3], dy)));
              const float yr = __fmaf_rn(r[q][7], dz, __fmaf_rn(r[q][1], dx, __fmul_rn(r[q][4], dy)));
              const float zr = __fmaf_rn(r[q][8], dz, __fmaf_rn(r[q][2], dx, __fmul_rn(r[q][5], dy)));
              const float d2 = __fmaf_rn(yr, yr, __fmul_rn(zr, zr));
              hit = valid && (d2 < radius2) && (xr > hmin) && (xr < hmax);
            } else {
              const float d2 = sqdist3(qx[q] - x, qy[q] - y, qz[q] - z);
              hit = valid && (d2 < radius2);
            }
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
    if (q0 + q < m) {
      const int c = min(cnt[q], nsample);
      const int fill = c > 0 ? first[q] : 0;
      for (int s = c + lane; s < nsample; s += 32) out[q][s] = fill;
    }
  }
}

template <bool CYL>
static int launch_query(const float *new_xyz, const float *xyz, const float *rot, int *idx, int b, int n, int m, float radius,
                        float hmin, float hmax, int nsample, cudaStream_t s) {
  if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || !new_xyz || !xyz || !idx || (CYL && !rot)) return (int)cudaErrorInvalidValue;
  if (b == 0 || m == 0) return 0;
  const float radius2 = radius * radius;  // fp32 product, as ball_query_gpu.cu:22
  const int use_bulk = (n % 4 == 0) && (((uintptr_t)xyz & 15u) == 0);
  int qpw = g_tuning.query_qpw;
  if (qpw != 1 && qpw != 2 && qpw != 4) {
    const long warps = (long)b * m;
    qpw = warps >= 4L * 16 * num_sms() ? 4 : (warps >= 2L * 16 * num_sms() ? 2 : 1);
  }
  const int per_cta = kQueryWarps * qpw;
  dim3 grid((m + per_cta - 1) / per_cta, b);
  if (grid.y > 65535) return (int)cudaErrorInvalidValue;
  switch (qpw) {
    case 4: query_kernel<CYL, 4><<<grid, kQueryWarps * 32, 0, s>>>(new_xyz, xyz, rot, idx, n, m, radius2, hmin, hmax, nsample, use_bulk); break;
    case 2: query_kernel<CYL, 2><<<grid, kQueryWarps * 32, 0, s>>>(new_xyz, xyz, rot, idx, n, m, radius2, hmin, hmax, nsample, use_bulk); break;
    default: query_kernel<CYL, 1><<<grid, kQueryWarps * 32, 0, s>>>(new_xyz, xyz, rot, idx, n, m, radius2, hmin, hmax, nsample, use_bulk); break;
  }
  count_launch();
  return finish_launch();
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
