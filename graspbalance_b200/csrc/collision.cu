// collision.cu -- batched grasp x point gripper-volume occupancy test (fp64).
//
// Replaces the numpy body of ModelFreeCollisionDetector.detect (collision_detector.py:23-41,55), which materialises a
// [G,N,3] fp64 array of transformed points (491 MB at G=1024, N=20000) and ten [G,N] boolean masks on the host.  Here one
// WARP owns a grasp (translation, rotation and the ten half-space thresholds live in registers), the scene is staged tile
// by tile in shared memory, every lane transforms one point per step and keeps six integer counters that are warp-reduced
// at the end.  Nothing but the six counts per grasp ever leaves the SM.
//
// Arithmetic (SURVEY.md A.7 and DESIGN.md "collision rounding"): d = p - T in fp64; t_j = fma(d2,R[2][j], fma(d1,R[1][j],
// d0*R[0][j])) -- the evaluation order of the OpenBLAS dgemm kernel numpy.matmul dispatches to (bit-identical on 9.6e5
// samples); the thresholds are computed on the host with the reference's own numpy expressions, so every compare sees
// the same two doubles as the reference.
#include "common.cuh"

namespace gb {

constexpr int kColWarps = 4;
constexpr int kColTile = 1024;  // points per tile: 24 KB

__global__ void __launch_bounds__(kColWarps * 32) collision_kernel(const double *__restrict__ points, int np, const double *__restrict__ T,
                                                                   const double *__restrict__ R, const double *__restrict__ thr, int g,
                                                                   unsigned long long *__restrict__ counts, int pts_per_split,
                                                                   const long long *__restrict__ scene_off) {
  __shared__ double tile[kColTile * 3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (scene_off) {  // batched: blockIdx.z = scene; its points are rows scene_off[z] .. scene_off[z+1] of the packed array
    const long long o0 = scene_off[blockIdx.z];
    points += 3 * o0;
    np = (int)(scene_off[blockIdx.z + 1] - o0);
    const size_t go = (size_t)blockIdx.z * g;
    T += go * 3, R += go * 9, thr += go * 10, counts += go * 6;
  }
  const int gi = blockIdx.x * kColWarps + warp;
  const bool gok = gi < g;
  const size_t gs = gok ? gi : 0;
  const double t0 = T[gs * 3], t1 = T[gs * 3 + 1], t2 = T[gs * 3 + 2];
  double r[9], h[10];
#pragma unroll
  for (int e = 0; e < 9; ++e) r[e] = R[gs * 9 + e];
#pragma unroll
  for (int e = 0; e < 10; ++e) h[e] = thr[gs * 10 + e];

  const int p_begin = blockIdx.y * pts_per_split;
  const int p_end = min(np, p_begin + pts_per_split);
  int cg = 0, cl = 0, cr = 0, cb = 0, cs = 0, ci = 0;
  for (int base = p_begin; base < p_end; base += kColTile) {
    const int tc = min(kColTile, p_end - base);
    __syncthreads();
    for (int e = tid; e < tc * 3; e += kColWarps * 32) tile[e] = points[(size_t)base * 3 + e];
    __syncthreads();
    if (!gok) continue;
    for (int off = 0; off < tc; off += 32) {
      const int e = off + lane;
      const bool valid = e < tc;
      const int es = valid ? e : 0;
      const double d0 = tile[es * 3] - t0, d1 = tile[es * 3 + 1] - t1, d2 = tile[es * 3 + 2] - t2;
      const double tz = __fma_rn(d2, r[8], __fma_rn(d1, r[5], __dmul_rn(d0, r[2])));
      const bool m1 = valid && (tz > h[0]) && (tz < h[1]);
      if (!__any_sync(0xffffffffu, m1)) continue;  // every mask needs m1
      const double tx = __fma_rn(d2, r[6], __fma_rn(d1, r[3], __dmul_rn(d0, r[0])));
      const double ty = __fma_rn(d2, r[7], __fma_rn(d1, r[4], __dmul_rn(d0, r[1])));
      const bool m2 = (tx > h[2]) && (tx < h[3]);
      const bool m3 = ty > h[4];
      const bool m4 = ty < h[5];
      const bool m5 = ty < h[6];
      const bool m6 = ty > h[7];
      const bool m7 = (tx <= h[2]) && (tx > h[8]);
      const bool m8 = (tx <= h[8]) && (tx > h[9]);
      const bool left = m1 && m2 && m3 && m4, right = m1 && m2 && m5 && m6;
      const bool bottom = m1 && m3 && m5 && m7, shifting = m1 && m3 && m5 && m8;
      cg += (left || right || bottom || shifting) ? 1 : 0;
      cl += left ? 1 : 0;
      cr += right ? 1 : 0;
      cb += bottom ? 1 : 0;
      cs += shifting ? 1 : 0;
      ci += (m1 && m2 && !m4 && !m6) ? 1 : 0;
    }
  }
  if (!gok) return;
  cg = __reduce_add_sync(0xffffffffu, cg);
  cl = __reduce_add_sync(0xffffffffu, cl);
  cr = __reduce_add_sync(0xffffffffu, cr);
  cb = __reduce_add_sync(0xffffffffu, cb);
  cs = __reduce_add_sync(0xffffffffu, cs);
  ci = __reduce_add_sync(0xffffffffu, ci);
  if (lane < 6) {
    const int v = lane == 0 ? cg : lane == 1 ? cl : lane == 2 ? cr : lane == 3 ? cb : lane == 4 ? cs : ci;
    if (v) atomicAdd(counts + (size_t)gi * 6 + lane, (unsigned long long)v);
  }
}

// Per-voxel means for the voxel down-sampling of ModelFreeCollisionDetector.__init__ (collision_detector.py:11-14, open3d's
// PointCloud.voxel_down_sample): the points of a voxel, listed in INPUT order, are summed sequentially in fp64 by one
// thread and divided by their count -- the running sum open3d keeps per voxel, so the means are bit-identical to the host
// restatement.  order [n] = point indices grouped by voxel (stable sort by voxel key), seg [v+1] = first position of
// every voxel in `order`.
__global__ void voxel_means_kernel(const double *__restrict__ points, const long long *__restrict__ order,
                                   const long long *__restrict__ seg, double *__restrict__ out, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v) return;
  const long long s0 = seg[i], s1 = seg[i + 1];
  double sx = 0.0, sy = 0.0, sz = 0.0;
  for (long long k = s0; k < s1; ++k) {
    const double *p = points + 3 * order[k];
    sx = __dadd_rn(sx, p[0]), sy = __dadd_rn(sy, p[1]), sz = __dadd_rn(sz, p[2]);
  }
  const double cnt = (double)(s1 - s0);
  out[3 * (size_t)i] = sx / cnt, out[3 * (size_t)i + 1] = sy / cnt, out[3 * (size_t)i + 2] = sz / cnt;
}

static int collision_launch(const double *points, int np, const double *T, const double *R, const double *thr, int g, int64_t *counts,
                            cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)g * 6 * sizeof(int64_t), s);
  if (e != cudaSuccess) return (int)e;
  if (np == 0) return 0;
  const int gx = (g + kColWarps - 1) / kColWarps;
  int splits = (4 * num_sms() + gx - 1) / gx;  // aim at ~4 CTAs per SM
  const int max_splits = (np + kColTile - 1) / kColTile;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int pps = (np + splits - 1) / splits;
  pps = ((pps + kColTile - 1) / kColTile) * kColTile;
  splits = (np + pps - 1) / pps;
  dim3 grid(gx, splits);
  collision_kernel<<<grid, kColWarps * 32, 0, s>>>(points, np, T, R, thr, g, reinterpret_cast<unsigned long long *>(counts), pps, nullptr);
  count_launch();
  return finish_launch();
}

}  // namespace gb

extern "C" int gb_collision_counts(const double *points, int np, const double *T, const double *R, const double *thr, int g,
                                   int64_t *counts, gb_stream_t stream) {
  if (np < 0 || g < 0 || !T || !R || !thr || !counts || (np > 0 && !points)) return (int)cudaErrorInvalidValue;
  if (g == 0) return 0;
  return gb::collision_launch(points, np, T, R, thr, g, counts, (cudaStream_t)stream);
}

extern "C" int gb_collision_counts_host(const double *points, int np, const double *T, const double *R, const double *thr, int g,
                                        int64_t *counts) {
  if (np < 0 || g < 0 || !T || !R || !thr || !counts || (np > 0 && !points)) return (int)cudaErrorInvalidValue;
  if (g == 0) return 0;
  cudaStream_t s = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  if (e != cudaSuccess) return (int)e;
  const size_t bp = (size_t)np * 3 * sizeof(double), bt = (size_t)g * 3 * sizeof(double), br = (size_t)g * 9 * sizeof(double),
               bh = (size_t)g * 10 * sizeof(double), bc = (size_t)g * 6 * sizeof(int64_t);
  unsigned char *d = nullptr;
  const size_t total = ((bp + 255) & ~(size_t)255) + ((bt + 255) & ~(size_t)255) + ((br + 255) & ~(size_t)255) + ((bh + 255) & ~(size_t)255) + bc;
  e = cudaMalloc((void **)&d, total);
  if (e != cudaSuccess) { cudaStreamDestroy(s); return (int)e; }
  unsigned char *dp = d, *dt = dp + ((bp + 255) & ~(size_t)255), *dr = dt + ((bt + 255) & ~(size_t)255),
                *dh = dr + ((br + 255) & ~(size_t)255), *dc = dh + ((bh + 255) & ~(size_t)255);
  int rc = 0;
  if (bp) rc = (int)cudaMemcpyAsync(dp, points, bp, cudaMemcpyHostToDevice, s);
  if (!rc) rc = (int)cudaMemcpyAsync(dt, T, bt, cudaMemcpyHostToDevice, s);
  if (!rc) rc = (int)cudaMemcpyAsync(dr, R, br, cudaMemcpyHostToDevice, s);
  if (!rc) rc = (int)cudaMemcpyAsync(dh, thr, bh, cudaMemcpyHostToDevice, s);
  if (!rc) rc = gb::collision_launch((const double *)dp, np, (const double *)dt, (const double *)dr, (const double *)dh, g, (int64_t *)dc, s);
  if (!rc) rc = (int)cudaMemcpyAsync(counts, dc, bc, cudaMemcpyDeviceToHost, s);
  if (!rc) rc = (int)cudaStreamSynchronize(s);
  cudaFree(d);
  cudaStreamDestroy(s);
  return rc;
}

/* Voxel means for the down-sampling step of ModelFreeCollisionDetector.__init__ (collision_detector.py:11-14): points [n,3]
 * f64, order [n] i64 = point indices grouped by voxel with the input order kept inside a voxel, seg [v+1] i64 = first
 * position of every voxel in `order`; out [v,3] f64 = sequential fp64 sum of the voxel's points divided by their count. */
extern "C" int gb_voxel_means(const double *points, const long long *order, const long long *seg, double *out, int v,
                              gb_stream_t stream) {
  if (v < 0) return (int)cudaErrorInvalidValue;
  if (v == 0) return 0;
  if (!points || !order || !seg || !out) return (int)cudaErrorInvalidValue;
  gb::voxel_means_kernel<<<(v + 127) / 128, 128, 0, (cudaStream_t)stream>>>(points, order, seg, out, v);
  gb::count_launch();
  return gb::finish_launch();
}

/* The occupancy test of several scenes in one launch: scene z has the rows scene_off[z] .. scene_off[z+1] of the
 * packed `points` [sum N', 3] f64 (scene_off: nscenes + 1 i64 on the DEVICE) and g grasps, T [nscenes, g, 3], R [nscenes, g,
 * 3, 3], thr [nscenes, g, 10]; counts [nscenes, g, 6] i64.  max_np = the largest scene (host knowledge: sizes the grid).
 * Per scene the counts equal gb_collision_counts on that scene alone. */
extern "C" int gb_collision_counts_batched(const double *points, const long long *scene_off, int nscenes, int max_np, const double *T,
                                           const double *R, const double *thr, int g, int64_t *counts, gb_stream_t stream) {
  if (nscenes < 0 || max_np < 0 || g < 0) return (int)cudaErrorInvalidValue;
  if (nscenes == 0 || g == 0) return 0;
  if (!scene_off || !T || !R || !thr || !counts || (max_np > 0 && !points) || nscenes > 65535) return (int)cudaErrorInvalidValue;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)nscenes * g * 6 * sizeof(int64_t), s);
  if (e != cudaSuccess) return (int)e;
  if (max_np == 0) return 0;
  const int gx = (g + gb::kColWarps - 1) / gb::kColWarps;
  int splits = (4 * gb::num_sms() + gx * nscenes - 1) / (gx * nscenes);
  const int max_splits = (max_np + gb::kColTile - 1) / gb::kColTile;
  splits = splits > max_splits ? max_splits : (splits < 1 ? 1 : splits);
  int pps = (max_np + splits - 1) / splits;
  pps = ((pps + gb::kColTile - 1) / gb::kColTile) * gb::kColTile;
  splits = (max_np + pps - 1) / pps;
  dim3 grid(gx, splits, nscenes);
  gb::collision_kernel<<<grid, gb::kColWarps * 32, 0, s>>>(points, 0, T, R, thr, g, reinterpret_cast<unsigned long long *>(counts), pps,
                                                           scene_off);
  gb::count_launch();
  return gb::finish_launch();
}
